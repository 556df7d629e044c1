/* Test infrastructure: the few declarations of libao's <ao/ao.h> the reference CLI (reference src/main.c:16,
 * 88-103,670,1000,1016,1137) needs.  libao is not installed in this image; tests/cli/stubs/ao_stub.c implements them
 * as a sink that counts the bytes played, so that the UNMODIFIED src/main.c can be compiled and linked against the
 * drop-in libnrsc5.so and the reference's own CI test (.github/workflows/ci.yml:30-42) run on it. */
#ifndef AO_STUB_H
#define AO_STUB_H
#include <errno.h>      /* the real ao.h pulls these in, and src/main.c relies on it */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define AO_FMT_LITTLE 1
#define AO_FMT_BIG 2
#define AO_FMT_NATIVE 4

typedef struct ao_device ao_device;
typedef struct ao_option ao_option;
typedef struct ao_sample_format {
    int bits;
    int rate;
    int channels;
    int byte_format;
    char *matrix;
} ao_sample_format;

void ao_initialize(void);
void ao_shutdown(void);
int ao_default_driver_id(void);
int ao_driver_id(const char *short_name);
ao_device *ao_open_live(int driver_id, ao_sample_format *format, ao_option *option);
ao_device *ao_open_file(int driver_id, const char *filename, int overwrite, ao_sample_format *format, ao_option *option);
int ao_play(ao_device *device, char *output_samples, uint32_t num_bytes);
int ao_close(ao_device *device);
#endif
