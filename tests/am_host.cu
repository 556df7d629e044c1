// Host harness for nrsc5_b200/csrc/am.cuh (test infrastructure): the AM engine's AM_HD functions compiled for the
// CPU and run with ONE lane, so that tests/test_am_host.py can hold them against the oracle without a GPU.
// Built by tests/test_am_host.py:  nvcc -shared -Xcompiler -fPIC -o tests/_build/libam_host.so tests/am_host.cu
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../nrsc5_b200/csrc/am.cuh"
#include "../nrsc5_b200/csrc/am_tables.h"

extern "C" int orc_fix_header(uint8_t *buf);     // oracle/nrsc5_oracle.c (test infrastructure may use the oracle)

extern "C" long am_host_decode(const int16_t *cs16, size_t nvalues, uint8_t *log, size_t log_cap)
{
    using namespace nbam;
    AmTables *tb = new AmTables;
    am_fill_tables(*tb);
    AmWork *w = (AmWork *)calloc(1, sizeof(AmWork));
    AmState st;
    memset(&st, 0, sizeof(st));
    am_reset_state(st);
    AmIo io = { cs16, log, (unsigned)log_cap };
    st.in_avail = (long long)(nvalues / 2);
    const Lanes L = { 0, 1 };
    while (st.in_avail >= st.start + NACQ)
        process_window(st, *w, *tb, io, L, [](uint8_t *pdu) { return orc_fix_header(pdu); });
    long n = (long)st.log_len;
    free(w);
    delete tb;
    return st.log_overflow ? -1 : n;
}
