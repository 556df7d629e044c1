/* Test infrastructure: libao stand-in for the reference CLI (see ao/ao.h next to this file). */
#include <stdio.h>
#include <stdlib.h>
#include "ao/ao.h"

struct ao_device {
    FILE *f;
    unsigned long long bytes;
};

void ao_initialize(void) {}
void ao_shutdown(void) {}
int ao_default_driver_id(void) { return 0; }
int ao_driver_id(const char *short_name) { (void)short_name; return 1; }

ao_device *ao_open_live(int driver_id, ao_sample_format *format, ao_option *option)
{
    (void)driver_id; (void)format; (void)option;
    return (ao_device *)calloc(1, sizeof(ao_device));
}

ao_device *ao_open_file(int driver_id, const char *filename, int overwrite, ao_sample_format *format, ao_option *option)
{
    (void)driver_id; (void)overwrite; (void)format; (void)option;
    ao_device *d = (ao_device *)calloc(1, sizeof(ao_device));
    if (d) d->f = fopen(filename, "wb");
    return d;
}

int ao_play(ao_device *device, char *output_samples, uint32_t num_bytes)
{
    if (device->f) fwrite(output_samples, 1, num_bytes, device->f);
    device->bytes += num_bytes;
    return 1;
}

int ao_close(ao_device *device)
{
    if (device->f) fclose(device->f);
    free(device);
    return 1;
}
