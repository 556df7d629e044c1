// Host harness for nrsc5_b200/csrc/am.cuh (test infrastructure): the AM engine's AM_HD functions compiled for the
// CPU, so that tests/test_am_host.py can hold them against the oracle without a GPU.
//   am_host_decode        one lane.
//   am_host_decode_lanes  emulates the warp of k_am (engine.cu): one fibre per lane with a private AmState (the
//                         kernel's per-thread copy) over the shared AmWork; AM_SYNC() is a barrier at which the
//                         scheduler switches lanes.  Lanes run one after the other up to their next barrier, in
//                         ascending or descending lane order, so a missing AM_SYNC() between a phase that writes
//                         and a phase that reads other lanes' results shows up as a wrong answer (a harsher
//                         schedule than the GPU's lock step); lanes arriving at different AM_SYNC() sites are
//                         reported as divergence.
// Built by tests/test_am_host.py:  nvcc -shared -Xcompiler -fPIC -o tests/_build/libam_host.so tests/am_host.cu
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static ucontext_t g_sched_ctx;
static ucontext_t *g_lane_ctx = nullptr;     // non-null while fibres run
static int g_cur_lane = 0, g_sync_site = 0;
static inline void am_host_sync(int site)
{
    if (!g_lane_ctx) return;
    g_sync_site = site;
    swapcontext(&g_lane_ctx[g_cur_lane], &g_sched_ctx);
}
#define AM_HOST_SYNC() am_host_sync(__LINE__)

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

namespace {
struct LaneJob {
    nbam::AmState st;          // every lane's private copy, like the registers / local memory of k_am
    nbam::AmWork *w;
    const nbam::AmTables *tb;
    nbam::AmIo io;
    nbam::Lanes L;
    bool done;
};
LaneJob *g_jobs = nullptr;

void lane_main()
{
    using namespace nbam;
    LaneJob &j = g_jobs[g_cur_lane];
    while (j.st.in_avail >= j.st.start + NACQ) {
        process_window(j.st, *j.w, *j.tb, j.io, j.L, [](uint8_t *pdu) { return orc_fix_header(pdu); });
        am_host_sync(-1);      // k_am: __syncwarp() between windows
    }
    j.done = true;
    g_sync_site = -2;
    swapcontext(&g_lane_ctx[g_cur_lane], &g_sched_ctx);
}
}   // namespace

// order: +1 = lanes scheduled 0..n-1, -1 = n-1..0.  Returns the log length, -1 log overflow, -2 lane states differ,
// -3 lanes met at different AM_SYNC() sites.
extern "C" long am_host_decode_lanes(const int16_t *cs16, size_t nvalues, uint8_t *log, size_t log_cap, int nlanes, int order)
{
    using namespace nbam;
    AmTables *tb = new AmTables;
    am_fill_tables(*tb);
    AmWork *w = (AmWork *)calloc(1, sizeof(AmWork));
    AmState st;
    memset(&st, 0, sizeof(st));
    am_reset_state(st);
    st.in_avail = (long long)(nvalues / 2);
    std::vector<LaneJob> jobs(nlanes);
    std::vector<ucontext_t> ctx(nlanes);
    std::vector<std::vector<char>> stacks(nlanes, std::vector<char>(1 << 20));
    g_jobs = jobs.data();
    for (int l = 0; l < nlanes; l++) {
        jobs[l] = LaneJob{ st, w, tb, AmIo{ cs16, log, (unsigned)log_cap }, Lanes{ l, nlanes }, false };
        getcontext(&ctx[l]);
        ctx[l].uc_stack.ss_sp = stacks[l].data();
        ctx[l].uc_stack.ss_size = stacks[l].size();
        ctx[l].uc_link = &g_sched_ctx;
        makecontext(&ctx[l], lane_main, 0);
    }
    g_lane_ctx = ctx.data();
    bool diverged = false;
    for (;;) {
        int site = 0, alive = 0;
        for (int k = 0; k < nlanes; k++) {
            const int l = order > 0 ? k : nlanes - 1 - k;
            if (jobs[l].done) continue;
            g_cur_lane = l;
            swapcontext(&g_sched_ctx, &ctx[l]);
            if (alive++ == 0) site = g_sync_site;
            else if (site != g_sync_site) diverged = true;
        }
        if (!alive || diverged) break;
    }
    g_lane_ctx = nullptr;
    g_jobs = nullptr;
    // every lane must have followed the same control flow and hold the same state (lane 0's is what k_am keeps)
    long n = (long)jobs[0].st.log_len;
    for (int l = 1; l < nlanes; l++)
        if (memcmp(&jobs[l].st, &jobs[0].st, sizeof(AmState)) != 0) n = -2;
    if (jobs[0].st.log_overflow) n = -1;
    if (diverged) n = -3;
    free(w);
    delete tb;
    return n;
}

// ---- the cu8 front end (decim_tile) under the same fibre emulation, with the engine's ring + tiling arithmetic ----
namespace {
struct DecimJob {
    const uint8_t *ring;
    unsigned ring_bytes;
    long long raw_avail, k0;
    int nout;
    short2 *out;
    nbam::DecimScratch *sc;
    int nlanes;
};
DecimJob g_dj;
bool g_dj_done[1024];

void decim_lane_main()
{
    const int l = g_cur_lane;
    nbam::decim_tile(g_dj.ring, g_dj.ring_bytes, g_dj.raw_avail, g_dj.k0, g_dj.nout, g_dj.out, *g_dj.sc, nbam::Lanes{ l, g_dj.nlanes });
    g_dj_done[l] = true;
    g_sync_site = -2;
    swapcontext(&g_lane_ctx[l], &g_sched_ctx);
}

// one tile on nlanes fibres (nlanes == 1: plain call); returns false when lanes met at different barriers
bool run_decim_tile(int nlanes, int order)
{
    if (nlanes == 1) {
        nbam::decim_tile(g_dj.ring, g_dj.ring_bytes, g_dj.raw_avail, g_dj.k0, g_dj.nout, g_dj.out, *g_dj.sc, nbam::Lanes{ 0, 1 });
        return true;
    }
    static std::vector<ucontext_t> ctx;
    static std::vector<std::vector<char>> stacks;
    if ((int)ctx.size() < nlanes) {
        ctx.resize(nlanes);
        stacks.assign(nlanes, std::vector<char>(256 << 10));
    }
    for (int l = 0; l < nlanes; l++) {
        g_dj_done[l] = false;
        getcontext(&ctx[l]);
        ctx[l].uc_stack.ss_sp = stacks[l].data();
        ctx[l].uc_stack.ss_size = stacks[l].size();
        ctx[l].uc_link = &g_sched_ctx;
        makecontext(&ctx[l], decim_lane_main, 0);
    }
    g_lane_ctx = ctx.data();
    bool ok = true;
    for (;;) {
        int site = 0, alive = 0;
        for (int k = 0; k < nlanes; k++) {
            const int l = order > 0 ? k : nlanes - 1 - k;
            if (g_dj_done[l]) continue;
            g_cur_lane = l;
            swapcontext(&g_sched_ctx, &ctx[l]);
            if (alive++ == 0) site = g_sync_site;
            else if (site != g_sync_site) ok = false;
        }
        if (!alive || !ok) break;
    }
    g_lane_ctx = nullptr;
    return ok;
}
}   // namespace

// cu8 -> cs16 the way the engine does it: pushes of `chunk` bytes land in a ring of `ring_bytes`, every push
// is followed by the tiles of the outputs that became computable (nrsc5b_push_cu8 in AM mode, engine.cu).
// Returns the number of cs16 complex samples written, or -3 on barrier divergence.
extern "C" long am_host_decimate(const uint8_t *cu8, size_t nbytes, int16_t *out, unsigned ring_bytes, size_t chunk, int nlanes,
                                 int order)
{
    using namespace nbam;
    std::vector<uint8_t> ring(ring_bytes);
    DecimScratch *sc = new DecimScratch;
    long long raw_bytes = 0, done = 0;
    nbytes &= ~(size_t)3;
    for (size_t off = 0; off < nbytes; off += chunk) {
        const size_t n = nbytes - off < chunk ? nbytes - off : chunk;
        for (size_t i = 0; i < n; i++) ring[(raw_bytes + (long long)i) & (ring_bytes - 1)] = cu8[off + i];
        raw_bytes += (long long)n;
        const long long raw_avail = raw_bytes / 2, can = raw_avail / 32;
        for (long long k0 = done; k0 < can; k0 += DEC_T) {
            const int nout = (int)(can - k0 < DEC_T ? can - k0 : DEC_T);
            g_dj = DecimJob{ ring.data(), ring_bytes, raw_avail, k0, nout, reinterpret_cast<short2 *>(out) + k0, sc, nlanes };
            if (!run_decim_tile(nlanes, order)) {
                delete sc;
                return -3;
            }
        }
        done = can;
    }
    delete sc;
    return (long)done;
}
