/*
 * amp_probe - links the UNMODIFIED reference air.c against a stand-in deqframe()
 * that writes every envelope sample to a file.  Used only to
 * pin oracle.c's front end bit-for-bit.  TEST INFRASTRUCTURE ONLY.
 *
 * usage: amp_probe capture.bin amp_out.f32
 */
#include <stdio.h>
#include "adsbdec.h"

extern char *filename;                 /* air.c:217 */
extern void *fileInput(void *arg);     /* air.c:219 */

static FILE *out;

/* Consumes everything but the DECOFFSET look-ahead, like a scan that never
 * accepts a frame (demod.c:89, 143). */
int deqframe(const float *ampbuff, const int len)
{
	int used = len - DECOFFSET;
	fwrite(ampbuff, sizeof(float), (size_t)used, out);
	return used;
}

void handlerExit(int sig) { (void)sig; }

int main(int argc, char **argv)
{
	if (argc != 3)
		return 2;
	filename = argv[1];
	out = fopen(argv[2], "wb");
	if (!out)
		return 1;
	fileInput(NULL);
	fclose(out);
	return 0;
}
