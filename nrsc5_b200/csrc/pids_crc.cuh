// CRC-12 of an 80-bit PIDS frame (reference src/pids.c:52-86, 1032-1050): the first thing the reference's L2 does with
// a PIDS PDU.  The engine computes it next to the decode and appends the verdict to the PIDS record so that a batch
// caller can drop bad frames without touching the bits.  `pk` = the 80 frame bits packed MSB-first,
// as in the record.  pids_frame_push first reverses the bits of every byte: its bit i is bit (i & 7) of byte i >> 3.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) || defined(NB_EMU)
#define NB_HD __host__ __device__
#else
#define NB_HD
#endif

NB_HD inline int pids_crc12_ok(const uint8_t *pk)
{
    auto bit = [&](int i) -> unsigned { return (pk[i >> 3] >> (i & 7)) & 1u; };
    const unsigned poly = 0xD010;
    unsigned reg = 0;
    for (int i = 67; i >= 0; i--) {
        const unsigned low = reg & 1;
        reg >>= 1;
        reg ^= bit(i) << 15;
        if (low) reg ^= poly;
    }
    for (int i = 0; i < 16; i++) {
        const unsigned low = reg & 1;
        reg >>= 1;
        if (low) reg ^= poly;
    }
    reg = (reg ^ 0x955) & 0xfff;
    unsigned expected = 0;
    for (int i = 68; i < 80; i++) expected = (expected << 1) | bit(i);
    return expected == reg;
}
