// format.cpp - packet writers with the observable behaviour of formatpkt() (output.c:204-262,
// WITH_AIR build): avr, avrmlat and beast.  Host code; nothing here touches the GPU.
#include <math.h>
#include <stdint.h>

#include "../../include/adsbdec_b200.h"

namespace {

const char kHex[] = "0123456789ABCDEF";

inline char *put_hex_byte(char *p, uint8_t b)
{
	*p++ = kHex[b >> 4];
	*p++ = kHex[b & 15];
	return p;
}

}  // namespace

extern "C" int adsbdec_b200_format_packet(int fmt, const uint8_t *frame, int len, uint64_t ts,
                                          uint32_t pw, char *pkt)
{
	if (!frame || !pkt || len < 0 || len > 112)
		return ADSB_B200_EINVAL;
	// output.c:218: the 10 MS/s iteration counter rescaled to the 12 MHz beast/mlat clock
	const uint64_t ts12 = ts * 12 / 10;
	char *p = pkt;
	if (fmt >= 2) {
		// output.c:215-217: the level is computed in double from the integer pw, so it is exact
		// on any IEEE host: nearbyint(sqrt(pw)) / 8 / 5, converted to a byte.
		const uint8_t lvl = (uint8_t)(nearbyint(sqrt((double)pw)) / 8 / 5);
		*p++ = 0x1a;
		*p++ = len == 7 ? '2' : '3';  // output.c:233
		for (int sh = 40; sh >= 0; sh -= 8) {
			const char ch = (char)(ts12 >> sh);
			*p++ = ch;
			if (ch == 0x1a)
				*p++ = ch;  // output.c:235-238: escape
		}
		*p++ = (char)lvl;  // output.c:241: the level byte is NOT escaped
		for (int i = 0; i < len; i++) {
			*p++ = (char)frame[i];
			if (frame[i] == 0x1a)
				*p++ = 0x1a;
		}
		return (int)(p - pkt);
	}
	if (fmt == 1) {
		*p++ = '@';  // output.c:226: "@%012" PRIX64 of the low 48 bits
		for (int sh = 44; sh >= 0; sh -= 4)
			*p++ = kHex[(ts12 >> sh) & 15];
	} else {
		*p++ = '*';  // output.c:222
	}
	for (int i = 0; i < len; i++)
		p = put_hex_byte(p, frame[i]);
	*p++ = ';';
	*p++ = '\n';
	*p = 0;
	return (int)(p - pkt);
}

// Every frame of an array through format_packet(), packets back to back: what runOutput() (output.c:286-336)
// writes for them.  Returns the bytes written, or ADSB_B200_ERANGE if `cap` is too small (a packet is at most
// 2 + 12 + 1 + 28 bytes binary, 13 + 28 + 3 as text).
extern "C" int64_t adsbdec_b200_format_frames(int fmt, const adsbdec_b200_frame_t *frames, size_t count, char *out,
                                              size_t cap)
{
	if ((!frames && count) || !out)
		return ADSB_B200_EINVAL;
	char pkt[256];
	size_t at = 0;
	for (size_t i = 0; i < count; i++) {
		const int n = adsbdec_b200_format_packet(fmt, frames[i].bytes, frames[i].len, frames[i].ts, frames[i].pw, pkt);
		if (n < 0)
			return n;
		if (at + (size_t)n > cap)
			return ADSB_B200_ERANGE;
		for (int k = 0; k < n; k++)
			out[at + k] = pkt[k];
		at += (size_t)n;
	}
	return (int64_t)at;
}
