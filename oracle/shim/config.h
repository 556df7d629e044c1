/* Hand-written stand-in for the reference's generated config.h
 * (template: reference src/config.h.in:1-25).  Test infrastructure only:
 * used solely to compile the reference sources into oracle/_ref/. */
#pragma once
#define HAVE_STRNDUP
#define HAVE_CMPLXF
#define HAVE_IMAGINARY_I
#define HAVE_COMPLEX_I
#define LIBRARY_DEBUG_LEVEL 5
