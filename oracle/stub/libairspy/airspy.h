/*
 * Minimal stand-in for <libairspy/airspy.h>, written for this repo so that the
 * UNMODIFIED reference sources under /root/reference compile without the SDR
 * library.  Only the names air.c:104-215 touches are
 * declared.  The `-f FILE` path never reaches any of them: airspy_open() in
 * airspy_stub.c fails, and startSDR() returns early on a NULL device.
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 */
#ifndef ORACLE_AIRSPY_STUB_H
#define ORACLE_AIRSPY_STUB_H
#include <stdint.h>

struct airspy_device;

enum { AIRSPY_SUCCESS = 0, AIRSPY_TRUE = 1, AIRSPY_ERROR_NOT_FOUND = -5 };
enum airspy_sample_type { AIRSPY_SAMPLE_FLOAT32_IQ = 0, AIRSPY_SAMPLE_RAW = 5 };

typedef struct {
	struct airspy_device *device;
	void *ctx;
	void *samples;
	int sample_count;
	uint64_t dropped_samples;
	enum airspy_sample_type sample_type;
} airspy_transfer_t;

typedef int (*airspy_sample_block_cb_fn)(airspy_transfer_t *transfer);

int airspy_open(struct airspy_device **device);
int airspy_close(struct airspy_device *device);
int airspy_set_sample_type(struct airspy_device *device, enum airspy_sample_type t);
int airspy_set_packing(struct airspy_device *device, uint8_t value);
int airspy_get_samplerates(struct airspy_device *device, uint32_t *buffer, const uint32_t len);
int airspy_set_samplerate(struct airspy_device *device, uint32_t samplerate);
int airspy_set_linearity_gain(struct airspy_device *device, uint8_t value);
int airspy_set_freq(struct airspy_device *device, const uint32_t freq_hz);
int airspy_start_rx(struct airspy_device *device, airspy_sample_block_cb_fn cb, void *ctx);
int airspy_stop_rx(struct airspy_device *device);
int airspy_is_streaming(struct airspy_device *device);
const char *airspy_error_name(int errcode);
#endif
