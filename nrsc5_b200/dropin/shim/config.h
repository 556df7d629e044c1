/* Hand-written stand-in for the reference's generated config.h (template: reference src/config.h.in:1-25),
 * used to compile the reference's unmodified host-side sources into the drop-in libnrsc5.so. */
#pragma once
#define HAVE_STRNDUP
#define HAVE_CMPLXF
#define HAVE_IMAGINARY_I
#define HAVE_COMPLEX_I
#define LIBRARY_DEBUG_LEVEL 5
