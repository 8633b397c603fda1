/*
 * oracle.c - CPU restatement of the adsbdec `-f` hot path.  TEST INFRASTRUCTURE ONLY.
 * See oracle.h for scope, citations and how this file is pinned.
 *
 * Build with -ffp-contract=off: every fused multiply-add below is explicit.
 */
#include "oracle.h"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

enum {
	PULSE = 5,                 /* adsbdec.h:1  PULSEW, 0.5 us at 10 MS/s */
	LOOKAHEAD = 240 * PULSE,   /* adsbdec.h:3  DECOFFSET */
	NTAPS = 14,                /* air.c:32     DSFLTLEN */
	AMP_FLUSH = 8196 * PULSE,  /* air.c:47     APBUFFSZ */
	READ_SAMPLES = 1024 * 1024 /* air.c:218    IQBUFFSZ */
};

/* air.c:36-45: one period of the triangular low-pass.  The reference stores it
 * twice back to back so that a rotated window can be read without a modulo. */
static const float taps[NTAPS] = {
	0.012627f, 0.025254f, 0.037881f, 0.050508f, 0.063135f, 0.075761f, 0.088388f,
	0.088388f, 0.075761f, 0.063135f, 0.050508f, 0.037881f, 0.025254f, 0.012627f,
};

struct oracle_ctx {
	int accept_df18;
	int fma_contract;
	int repair;                 /* EXTENSION, 0 = the reference: see try_repair() */
	int ap;                     /* EXTENSION, 0 = the reference: see oracle_set_address_parity() */
	int iq;                     /* EXTENSION, 0 = the reference: see oracle_set_iq() */
	uint64_t *first_seen;       /* [2^24], by ICAO address: where it was first announced (~0: never) */
	/* front end state (air.c:33-34, 49-50) */
	float ring[NTAPS];
	uint32_t ring_pos;
	float amp[AMP_FLUSH + 4];
	uint32_t amp_fill;
	uint64_t amp_base;          /* absolute index of amp[0] */
	/* demod state (demod.c:86) */
	uint64_t ts;
	/* stats (valid.c:30-31) */
	uint32_t tried[32], passed[32];
	/* results */
	oracle_frame_t *frames;
	size_t nframes, capframes;
	/* optional amp tap */
	float *tap_dst;
	size_t tap_cap, tap_n;
};

/* ---- CRC-24, crc.h:1-42 ------------------------------------------------- */

static uint32_t crc_tab[256];
static int crc_ready;

static void crc_build(void)
{
	/* crc.h:1-34 is the byte-wise table of the Mode S generator 0x1FFF409. */
	for (int b = 0; b < 256; b++) {
		uint32_t r = (uint32_t)b << 16;
		for (int k = 0; k < 8; k++)
			r = (r & 0x800000) ? ((r << 1) ^ 0xFFF409) : (r << 1);
		crc_tab[b] = r & 0xFFFFFF;
	}
	crc_ready = 1;
}

const uint32_t *oracle_crc_table(void)
{
	if (!crc_ready)
		crc_build();
	return crc_tab;
}

/* valid.c:48-50 / 70-72: CRC over all but the last three bytes, XOR the last
 * three in, keep 24 bits.  (crc.h:36-38 lets the register grow past 24 bits and
 * masks once at the end, crc.h:40-42; masking every step is the same value.) */
static uint32_t syndrome(const uint8_t *msg, int len)
{
	uint32_t r = 0;
	for (int i = 0; i < len - 3; i++)
		r = ((r << 8) ^ crc_tab[msg[i] ^ ((r >> 16) & 0xFF)]) & 0xFFFFFF;
	return r ^ ((uint32_t)msg[len - 3] << 16 | (uint32_t)msg[len - 2] << 8 | msg[len - 1]);
}

/* ---- EXTENSION: error repair (not in the reference) ---------
 * The reference stops at "syndrome != 0" (valid.c:52, :74).  With repair >= 1 a frame whose
 * syndrome is that of ONE flipped bit is corrected and accepted; with repair >= 2 also TWO
 * flipped bits of a 112-bit frame.  The five DF bits are never touched.  Stated here by brute
 * force over the candidate bit positions (the CUDA path uses a hash of syndromes instead), with
 * each position's syndrome taken from the table CRC above, so the two share no code. */
static int try_repair(const oracle_ctx *c, uint8_t *msg, int len)
{
	const int nbits = 8 * len;
	uint32_t s = syndrome(msg, len), one[112];
	for (int b = 0; b < nbits; b++) {
		uint8_t e[14] = {0};
		e[b >> 3] = (uint8_t)(0x80 >> (b & 7));
		one[b] = syndrome(e, len);
	}
	for (int b = 5; b < nbits; b++)
		if (one[b] == s) {
			msg[b >> 3] ^= (uint8_t)(0x80 >> (b & 7));
			return 1;
		}
	if (c->repair < 2 || len != 14)
		return 0;
	for (int b1 = 5; b1 < nbits; b1++)
		for (int b2 = b1 + 1; b2 < nbits; b2++)
			if ((one[b1] ^ one[b2]) == s) {
				msg[b1 >> 3] ^= (uint8_t)(0x80 >> (b1 & 7));
				msg[b2 >> 3] ^= (uint8_t)(0x80 >> (b2 & 7));
				return 2;
			}
	return 0;
}

/* ---- demodulator, demod.c ----------------------------------------------- */

static void emit(oracle_ctx *c, const uint8_t *msg, int len, uint32_t pw, uint64_t n, int fixed)
{
	if (c->nframes == c->capframes) {
		c->capframes = c->capframes ? 2 * c->capframes : 1024;
		c->frames = realloc(c->frames, c->capframes * sizeof *c->frames);
	}
	oracle_frame_t *f = &c->frames[c->nframes++];
	memset(f, 0, sizeof *f);
	f->ts = c->ts;
	f->n = n;
	f->pw = pw;
	f->len = len;
	memcpy(f->bytes, msg, (size_t)len);
	f->pad_[0] = (uint8_t)fixed;
}

/* demod.c:31-44: bit k of a byte is "pulse in first half-microsecond". */
static inline int ppm_bit(const float *a, int at)
{
	return a[at] > a[at + PULSE];
}

/* demod.c:84-144.  Returns how many samples were consumed. */
static int scan(oracle_ctx *c, const float *a, int have)
{
	int at = 0;
	while (at < have - LOOKAHEAD) {
		c->ts++;                                                   /* demod.c:99 */
		/* demod.c:102-105: float sums truncated to int */
		int p1 = (int)(a[at] + a[at + 2 * PULSE]);
		int s1 = (int)(a[at + PULSE] + a[at + 3 * PULSE]);
		int p2 = (int)(a[at + 7 * PULSE] + a[at + 9 * PULSE]);
		int s2 = (int)(a[at + 6 * PULSE] + a[at + 8 * PULSE]);
		if (!(p1 > 2 * s1 && p2 > 2 * s2)) {                       /* demod.c:107 */
			at++;
			continue;
		}
		/* demod.c:46-81: only 10001xxx, 10010xxx (-a) and 01011xxx survive */
		int d = at + 16 * PULSE;
		int b7 = ppm_bit(a, d), b6 = ppm_bit(a, d + 10), b5 = ppm_bit(a, d + 20);
		int b4 = ppm_bit(a, d + 30), b3 = ppm_bit(a, d + 40);
		int len = 0;
		if (b7 && !b6 && !b5) {
			if (b4 && !b3 && c->accept_df18)
				len = 14;       /* DF18 */
			else if (!b4 && b3)
				len = 14;       /* DF17 */
		} else if (!b7 && b6 && !b5 && b4 && b3) {
			len = 7;            /* DF11 */
		}
		int ap = 0;
		if (!len && c->ap && b5 && !b6 && !b4) {
			/* EXTENSION: 00100 / 00101 (DF4, DF5: 56 bits), 10100 / 10101 (DF20, DF21: 112 bits) */
			len = b7 ? 14 : 7;
			ap = 1;
		}
		if (!len) {
			at++;
			continue;
		}
		uint8_t msg[14];
		for (int i = 0; i < len; i++) {
			int v = 0;
			for (int k = 0; k < 8; k++)
				v = (v << 1) | ppm_bit(a, d + 80 * i + 10 * k);
			msg[i] = (uint8_t)v;
		}
		int df = msg[0] >> 3;
		c->tried[df]++;                                            /* valid.c:46, 68 */
		int fixed = 0;
		if (ap) {
			/* EXTENSION: the parity field carries CRC ^ address; the frame counts if that address was
			 * announced by a clean DF11/17/18 frame that starts earlier in the stream */
			uint32_t addr = syndrome(msg, len);
			if (addr != 0 && c->first_seen[addr] < c->amp_base + (uint64_t)at) {
				c->passed[df]++;
				emit(c, msg, len, (uint32_t)((p1 + p2) / 4), c->amp_base + (uint64_t)at, 0);
				at = d + 80 * len;
				continue;
			}
			at++;
			continue;
		}
		if (syndrome(msg, len) == 0 || (c->repair > 0 && (fixed = try_repair(c, msg, len)) > 0)) {
			c->passed[df]++;
			emit(c, msg, len, (uint32_t)((p1 + p2) / 4), c->amp_base + (uint64_t)at, fixed);
			at = d + 80 * len;                                     /* demod.c:128, 134 */
			continue;
		}
		at++;
	}
	return at;
}

/* ---- envelope front end, air.c:54-101 ----------------------------------- */

static inline void ring_put(oracle_ctx *c, float v)
{
	c->ring[c->ring_pos % NTAPS] = v;   /* air.c:65: uint32 counter, modulo 14 */
	c->ring_pos++;
}

static inline float envelope(const oracle_ctx *c)
{
	/* air.c:68-74: ascending ring slot; even slots feed I, odd slots feed Q;
	 * tap for slot k is taps[(k - ring_pos) mod 14]. */
	unsigned rot = NTAPS - c->ring_pos % NTAPS;
	float si = 0.0f, sq = 0.0f;
	for (int k = 0; k < NTAPS; k += 2) {
		float ti = taps[(k + rot) % NTAPS], tq = taps[(k + 1 + rot) % NTAPS];
		if (c->fma_contract) {
			si = fmaf(ti, c->ring[k], si);
			sq = fmaf(tq, c->ring[k + 1], sq);
		} else {
			si = si + ti * c->ring[k];
			sq = sq + tq * c->ring[k + 1];
		}
	}
	/* air.c:76; gcc -O3 on an FMA target rounds Q*Q, then fuses I*I + that */
	if (c->fma_contract) {
		float qq = sq * sq;
		return fmaf(si, si, qq);
	}
	float ii = si * si, qq = sq * sq;
	return ii + qq;
}

static inline void amp_put(oracle_ctx *c, float v)
{
	c->amp[c->amp_fill++] = v;
	if (c->tap_dst && c->tap_n < c->tap_cap)
		c->tap_dst[c->tap_n] = v;
	c->tap_n++;
}

void oracle_push(oracle_ctx *c, const uint16_t *buf, int len)
{
	if (c->iq) {
		/* EXTENSION: signed I/Q pairs, one per envelope sample, |z|^2 in the form of air.c:76; the
		 * buffer schedule (air.c:94-99: a look every two envelope samples) is the reference's */
		for (int i = 0; i < len; i += 4) {
			for (int k = 0; k < 4; k += 2) {
				float fi = (float)(int16_t)buf[i + k], fq = (float)(int16_t)buf[i + k + 1];
				amp_put(c, c->fma_contract ? fmaf(fi, fi, fq * fq) : fi * fi + fq * fq);
			}
			if (c->amp_fill >= AMP_FLUSH) {
				int used = scan(c, c->amp, (int)c->amp_fill);
				if ((uint32_t)used < c->amp_fill)
					memmove(c->amp, c->amp + used, (c->amp_fill - (uint32_t)used) * sizeof(float));
				c->amp_fill -= (uint32_t)used;
				c->amp_base += (uint64_t)used;
			}
		}
		return;
	}
	for (int i = 0; i < len; i += 4) {
		/* air.c:64-67, 79-82: centre on mid-scale; the second pair of every
		 * four is negated (fs/4 mixer). */
		ring_put(c, (float)buf[i] - 2048.0f);
		ring_put(c, (float)buf[i + 1] - 2048.0f);
		amp_put(c, envelope(c));
		ring_put(c, -((float)buf[i + 2] - 2048.0f));
		ring_put(c, -((float)buf[i + 3] - 2048.0f));
		amp_put(c, envelope(c));
		if (c->amp_fill >= AMP_FLUSH) {                            /* air.c:94-99 */
			int used = scan(c, c->amp, (int)c->amp_fill);
			if ((uint32_t)used < c->amp_fill)
				memmove(c->amp, c->amp + used, (c->amp_fill - (uint32_t)used) * sizeof(float));
			c->amp_fill -= (uint32_t)used;
			c->amp_base += (uint64_t)used;
		}
	}
}

/* ---- file / memory drivers, air.c:219-246 -------------------------------- */

void oracle_decode_memory(oracle_ctx *c, const void *data, size_t nbytes)
{
	/* one reused 2 MiB buffer; a ragged final read leaves earlier data behind it */
	uint16_t *buf = calloc(READ_SAMPLES + 4, sizeof *buf);
	const uint8_t *p = data;
	while (nbytes > 0) {
		size_t n = nbytes < READ_SAMPLES * sizeof *buf ? nbytes : READ_SAMPLES * sizeof *buf;
		memcpy(buf, p, n);
		oracle_push(c, buf, (int)(n / 2));
		p += n;
		nbytes -= n;
	}
	free(buf);
}

int oracle_decode_file(oracle_ctx *c, const char *path)
{
	int fd = open(path, O_RDONLY);
	if (fd < 0)
		return -1;
	uint16_t *buf = calloc(READ_SAMPLES + 4, sizeof *buf);
	for (;;) {
		ssize_t n = read(fd, buf, READ_SAMPLES * sizeof *buf);
		if (n <= 0)
			break;
		oracle_push(c, buf, (int)(n / 2));
	}
	free(buf);
	close(fd);
	return 0;
}

/* ---- bookkeeping --------------------------------------------------------- */

oracle_ctx *oracle_new(int accept_df18, int fma_contract)
{
	if (!crc_ready)
		crc_build();
	oracle_ctx *c = calloc(1, sizeof *c);
	c->accept_df18 = accept_df18;
	c->fma_contract = fma_contract;
	return c;
}

void oracle_set_repair(oracle_ctx *c, int level) { c->repair = level; }

/* EXTENSION (not in the reference, which only asks the airspy for real samples, air.c:128-152): the input is
 * complex baseband - signed 16-bit I/Q pairs at 10 MS/s, libairspy's INT16_IQ sample type, for which the library
 * does the fs/4 translation and the decimation itself.  One pair per envelope sample, envelope = I*I + Q*Q;
 * demod.c / valid.c behind it unchanged.  Integer sums as in demod.c:102-105: exact for |I|, |Q| <= 8191. */
void oracle_set_iq(oracle_ctx *c, int on) { c->iq = on; }

/* ---- EXTENSION: address/parity frames (not in the reference) ---------------------------
 * DF4, DF5 (56 bits) and DF20, DF21 (112 bits) carry CRC ^ ICAO address in their parity field, so their
 * syndrome is an address, not zero, and the reference (valid.c:52, :74) never accepts them.  With this mode on
 * such a frame is accepted when its address has been ANNOUNCED before it: by a DF11 / DF17 / (with -a) DF18
 * frame with syndrome 0 whose AA field (bytes 1..3) is that address and which starts at an earlier offset of
 * the stream.  "Starts at an offset" means: the preamble test, the DF gate and the CRC hold there - whether or
 * not the sequential scan visits that offset (it skips the inside of accepted frames).  That makes the set of
 * announcements a property of the capture alone, which a parallel decoder can establish without the scan.
 * Everything sequential (greedy skip, ts, tail) treats an accepted address/parity frame like any other; the
 * Try/Ok statistics (DF11/17/18) do not count them.  oracle_address_prepass() has to see the whole capture
 * before it is decoded. */
static int regular_candidate_at(const oracle_ctx *c, const float *a, uint8_t *msg)
{
	int p1 = (int)(a[0] + a[2 * PULSE]), s1 = (int)(a[PULSE] + a[3 * PULSE]);
	int p2 = (int)(a[7 * PULSE] + a[9 * PULSE]), s2 = (int)(a[6 * PULSE] + a[8 * PULSE]);
	if (!(p1 > 2 * s1 && p2 > 2 * s2))
		return 0;
	const int d = 16 * PULSE;
	int b7 = ppm_bit(a, d), b6 = ppm_bit(a, d + 10), b5 = ppm_bit(a, d + 20);
	int b4 = ppm_bit(a, d + 30), b3 = ppm_bit(a, d + 40);
	int len = 0;
	if (b7 && !b6 && !b5) {
		if (b4 && !b3 && c->accept_df18)
			len = 14;
		else if (!b4 && b3)
			len = 14;
	} else if (!b7 && b6 && !b5 && b4 && b3) {
		len = 7;
	}
	if (!len)
		return 0;
	for (int i = 0; i < len; i++) {
		int v = 0;
		for (int k = 0; k < 8; k++)
			v = (v << 1) | ppm_bit(a, d + 80 * i + 10 * k);
		msg[i] = (uint8_t)v;
	}
	return syndrome(msg, len) == 0 ? len : 0;
}

void oracle_set_address_parity(oracle_ctx *c, int on)
{
	c->ap = on;
	if (on && !c->first_seen) {
		c->first_seen = malloc(sizeof(uint64_t) << 24);
		memset(c->first_seen, 0xFF, sizeof(uint64_t) << 24);
	}
}

void oracle_address_prepass(oracle_ctx *c, const void *data, size_t nbytes)
{
	/* the envelope of the whole capture, through a second context with the same front-end state */
	size_t n_amp = nbytes / 4 * 2 + 8;
	float *amp = malloc(n_amp * sizeof *amp);
	oracle_ctx *t = oracle_new(c->accept_df18, c->fma_contract);
	t->ring_pos = c->ring_pos;
	t->iq = c->iq;
	oracle_record_amp(t, amp, n_amp);
	oracle_decode_memory(t, data, nbytes);
	size_t have = t->tap_n < n_amp ? t->tap_n : n_amp;
	oracle_free(t);
	uint8_t msg[14];
	for (size_t at = 0; at + LOOKAHEAD < have; at++)
		if (regular_candidate_at(c, amp + at, msg)) {
			uint32_t aa = (uint32_t)msg[1] << 16 | (uint32_t)msg[2] << 8 | msg[3];
			if (c->first_seen[aa] > (uint64_t)at)
				c->first_seen[aa] = (uint64_t)at;
		}
	free(amp);
}

void oracle_set_origin(oracle_ctx *c, uint64_t origin)
{
	c->ring_pos = (uint32_t)origin;
}

void oracle_free(oracle_ctx *c)
{
	if (c) {
		free(c->frames);
		free(c->first_seen);
		free(c);
	}
}

size_t oracle_num_frames(const oracle_ctx *c) { return c->nframes; }
const oracle_frame_t *oracle_frames(const oracle_ctx *c) { return c->frames; }

void oracle_stats(const oracle_ctx *c, uint32_t try_[3], uint32_t ok[3])
{
	static const int dfs[3] = { 11, 17, 18 };
	for (int i = 0; i < 3; i++) {
		try_[i] = c->tried[dfs[i]];
		ok[i] = c->passed[dfs[i]];
	}
}

void oracle_record_amp(oracle_ctx *c, float *dst, size_t capacity)
{
	c->tap_dst = dst;
	c->tap_cap = capacity;
	c->tap_n = 0;
}

size_t oracle_amp_count(const oracle_ctx *c) { return c->tap_n; }

/* ---- packet text, output.c:204-262 -------------------------------------- */

int oracle_format(const oracle_frame_t *f, int fmt, char *pkt)
{
	static const char hex[] = "0123456789ABCDEF";
	uint64_t ts12 = f->ts * 12 / 10;                               /* output.c:218 */
	char *p = pkt;
	if (fmt == 2) {
		/* output.c:215-217, WITH_AIR build: double sqrt -> nearbyint -> /8/5 -> uint8 */
		uint8_t lvl = (uint8_t)(nearbyint(sqrt((double)f->pw)) / 8 / 5);
		*p++ = 0x1a;
		*p++ = f->len == 7 ? '2' : '3';
		for (int sh = 40; sh >= 0; sh -= 8) {
			char ch = (char)(ts12 >> sh);
			*p++ = ch;
			if (ch == 0x1a)
				*p++ = ch;
		}
		*p++ = (char)lvl;                                          /* not escaped, output.c:241 */
		for (int i = 0; i < f->len; i++) {
			*p++ = (char)f->bytes[i];
			if (f->bytes[i] == 0x1a)
				*p++ = 0x1a;
		}
		return (int)(p - pkt);
	}
	if (fmt == 1) {
		*p++ = '@';
		for (int sh = 44; sh >= 0; sh -= 4)
			*p++ = hex[(ts12 >> sh) & 0xF];
	} else {
		*p++ = '*';
	}
	for (int i = 0; i < f->len; i++) {
		*p++ = hex[f->bytes[i] >> 4];
		*p++ = hex[f->bytes[i] & 0xF];
	}
	*p++ = ';';
	*p++ = '\n';
	*p = 0;
	return (int)(p - pkt);
}
