// Constant tables and the reset state of the AM engine (am.cuh); host code, shared by engine.cu and the CPU harness.
#pragma once
#include <cmath>
#include <cstring>

#include "am.cuh"

namespace nbam {

inline void am_fill_tables(AmTables &tb)
{
    // reference src/acquire.c:63-96 (band-pass taps), :333-342 (pulse shape)
    static const float coeff[32] = {
        -0.00038464731187559664f, -0.00021618751634377986f, 0.0026779419276863337f, -0.00029802651260979474f,
        -0.0012626448879018426f, -0.0013182522961869836f, -0.012252614833414555f, 0.015980124473571777f,
        0.037112727761268616f, -0.05451361835002899f, -0.05804193392395973f, 0.11320608854293823f,
        0.055298302322626114f, -0.16878043115139008f, -0.022917453199625015f, 0.19178225100040436f,
        -0.022917453199625015f, -0.16878043115139008f, 0.055298302322626114f, 0.11320608854293823f,
        -0.05804193392395973f, -0.05451361835002899f, 0.037112727761268616f, 0.015980124473571777f,
        -0.012252614833414555f, -0.0013182522961869836f, -0.0012626448879018426f, -0.00029802651260979474f,
        0.0026779419276863337f, -0.00021618751634377986f, -0.00038464731187559664f, 0.0f
    };
    for (int i = 0; i < 32; i++) tb.bp_tap[i] = (short)(coeff[31 - i] * 32767.0f);
    for (int i = 0; i < SYM; i++) {
        if (i < CP) tb.shape[i] = sinf((float)(M_PI / 2 * i / CP));
        else if (i < FFT) tb.shape[i] = 1;
        else tb.shape[i] = cosf((float)(M_PI / 2 * (i - FFT) / CP));
    }
    for (int k = 0; k < FFT / 2; k++) {
        const double a = -2.0 * M_PI * k / FFT;
        tb.tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int i = 0; i < FFT; i++) {
        unsigned r = 0;
        for (int b = 0; b < 8; b++) r |= ((i >> b) & 1u) << (7 - b);
        tb.brev[i] = (uint8_t)r;
    }
    unsigned reg = 0x3ff;                                   // reference src/decode.c:279-294
    for (int i = 0; i < P3_LEN_MA3 + 8; i++) {
        const unsigned b = ((reg >> 9) ^ reg) & 1;
        reg |= b << 11;
        reg >>= 1;
        tb.pn[i] = (uint8_t)b;
    }
}

inline void am_reset_state(AmState &st)                     // input_reset in AM mode (src/input.c:126-138)
{
    const long long avail = st.in_avail;
    memset(&st, 0, sizeof(st));
    st.in_avail = avail;
    st.phase = make_float2(1.0f, 0.0f);
    st.psmi = 1;
    st.pli = st.hppi = st.aabi = st.rdbi = -1;
    st.state = ST_NONE;
    st.am_diversity_wait = 4;
}

}  // namespace nbam
