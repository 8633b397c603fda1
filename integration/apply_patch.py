#!/usr/bin/env python
"""Applies the INTEGRATION.md section 3 patch to a checkout of the reference (TLeconte/adsbdec).

  python integration/apply_patch.py /root/reference oracle/_ref/patched

Writes patched copies of air.c and valid.c into the output directory (a build directory: nothing of the
reference is kept in this repository; the replacement text below is all there is):

  air.c    fileInput() (air.c:219-246) - open / 2 MiB reads / decodeiq() / handlerExit() - becomes five calls
           into libadsbdec_b200 (include/adsbdec_b200.h), with netout() (output.c:159) as the frame sink.
           Everything else in the file (the airspy callback path, decodeiq itself) stays.
  valid.c  gets a three-line setter so that print_stats() (valid.c:84-101) shows the library's Try/Ok counters
           (the arrays are file-static).

main.c, output.c, demod.c, crc.h and adsbdec.h are used as they are.  oracle/Makefile (`make patched`) builds
the result and tests/test_gpu_parity.py runs the golden captures through it.
"""
import re
import sys
from pathlib import Path

NEW_FILE_INPUT = r'''
/* ---- libadsbdec_b200: the GPU path behind the reference's -f input (INTEGRATION.md section 3) ---- */
#include "adsbdec_b200.h"
#include <stdio.h>
#include <unistd.h>

extern void netout(const uint8_t *frame, const int len, const uint64_t ts, const uint32_t pw); /* output.c:159 */
extern void set_stats(const uint32_t *tried, const uint32_t *ok);                              /* valid.c, patched */
extern int df;                                                                                  /* demod.c:26 */

static void b200_sink(const uint8_t *frame, int len, uint64_t ts, uint32_t pw, void *user)
{
	(void)user;
	netout(frame, len, ts, pw); /* copies the frame, wakes the output thread (output.c:159-182) */
}

void *fileInput(void *arg)
{
	adsbdec_b200_config_t cfg;
	adsbdec_b200_t *dec;
	uint32_t tried[3], ok[3];
	(void)arg;

	adsbdec_b200_default_config(&cfg);
	cfg.accept_df18 = df; /* -a */
	if (adsbdec_b200_open(&cfg, &dec) < 0) {
		fprintf(stderr, "%s\n", adsbdec_b200_last_error());
		handlerExit(0);
		return NULL;
	}
	adsbdec_b200_set_sink(dec, b200_sink, NULL);
	adsbdec_b200_decode_file(dec, filename); /* the 2 MiB read loop, decodeiq and EOF of the old fileInput */
	if (adsbdec_b200_stats(dec, tried, ok) == 0)
		set_stats(tried, ok);
	/* runOutput() looks at do_exit before its queue (output.c:286-288): let it drain what netout() queued */
	usleep(100000 + 25 * (useconds_t)adsbdec_b200_num_frames(dec));
	adsbdec_b200_close(dec);
	handlerExit(0);
	return NULL;
}
'''

STATS_SETTER = r'''
/* libadsbdec_b200 (INTEGRATION.md section 3): the library counts Try/Ok itself */
void set_stats(const uint32_t *tried, const uint32_t *ok)
{
	stat_try[11] = tried[0]; stat_try[17] = tried[1]; stat_try[18] = tried[2];
	stat_ok[11] = ok[0]; stat_ok[17] = ok[1]; stat_ok[18] = ok[2];
}
'''


def main():
    ref, out = Path(sys.argv[1]), Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    air = (ref / "air.c").read_text()
    m = re.search(r"^void \*fileInput\(void \*arg\)\s*\n\{.*?^\}\s*$", air, re.S | re.M)
    if not m:
        raise SystemExit("fileInput() not found in air.c: the reference changed")
    (out / "air.c").write_text(air[:m.start()] + NEW_FILE_INPUT.lstrip("\n") + air[m.end():])
    valid = (ref / "valid.c").read_text()
    if "stat_try[26]" not in valid:
        raise SystemExit("valid.c: stat_try not found: the reference changed")
    (out / "valid.c").write_text(valid + STATS_SETTER)
    print("patched air.c and valid.c ->", out)


if __name__ == "__main__":
    main()
