/* Stub bodies for oracle/stub/libairspy/airspy.h: every call reports "no device".
 * TEST INFRASTRUCTURE ONLY. */
#include <libairspy/airspy.h>
#include <stddef.h>

int airspy_open(struct airspy_device **device) { *device = NULL; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_close(struct airspy_device *device) { (void)device; return AIRSPY_SUCCESS; }
int airspy_set_sample_type(struct airspy_device *d, enum airspy_sample_type t) { (void)d; (void)t; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_set_packing(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_get_samplerates(struct airspy_device *d, uint32_t *b, const uint32_t len) { (void)d; (void)len; if (b) *b = 0; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_set_samplerate(struct airspy_device *d, uint32_t s) { (void)d; (void)s; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_set_linearity_gain(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_set_freq(struct airspy_device *d, const uint32_t f) { (void)d; (void)f; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_start_rx(struct airspy_device *d, airspy_sample_block_cb_fn cb, void *ctx) { (void)d; (void)cb; (void)ctx; return AIRSPY_ERROR_NOT_FOUND; }
int airspy_stop_rx(struct airspy_device *d) { (void)d; return AIRSPY_SUCCESS; }
int airspy_is_streaming(struct airspy_device *d) { (void)d; return 0; }
const char *airspy_error_name(int errcode) { (void)errcode; return "no airspy (stub)"; }
