/*
 * amp_probe - compiles the UNMODIFIED reference air.c into this translation unit (by #include,
 * from where it lies under /root/reference; nothing is copied) together with a stand-in
 * deqframe() that writes every envelope sample to a file.  Including the
 * source gives white-box access to its static `fidx` (air.c:34), so the 2^32 wrap
 * can be pinned without an 8.6 GB capture.  Used only to pin oracle.c's front end bit-for-bit.
 * TEST INFRASTRUCTURE ONLY.
 *
 * usage: amp_probe capture.bin amp_out.f32 [initial_fidx]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/types.h>
#include "air.c"

static FILE *out;

/* Consumes everything but the DECOFFSET look-ahead, like a scan that never
 * accepts a frame (demod.c:89, 143). */
int deqframe(const float *buf, const int len)
{
	int used = len - DECOFFSET;
	fwrite(buf, sizeof(float), (size_t)used, out);
	return used;
}

void handlerExit(int sig) { (void)sig; }

int main(int argc, char **argv)
{
	if (argc < 3)
		return 2;
	filename = argv[1];
	out = fopen(argv[2], "wb");
	if (!out)
		return 1;
	if (argc > 3)
		fidx = (uint32_t)strtoull(argv[3], NULL, 0);
	fileInput(NULL);
	fclose(out);
	return 0;
}
