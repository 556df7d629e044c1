// Fibre runtime of the CUDA emulator (see cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

#include <sys/mman.h>
#include <time.h>

#include <map>

#if !defined(__x86_64__)
#error "the emulator's context switch is written for x86-64"
#endif

// void emu_switch(void **save_sp, void *load_sp): saves the callee-saved registers of the System V ABI on the
// current stack, stores the stack pointer, continues on the other stack
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "    pushq %rbp\n"
    "    pushq %rbx\n"
    "    pushq %r12\n"
    "    pushq %r13\n"
    "    pushq %r14\n"
    "    pushq %r15\n"
    "    movq %rsp, (%rdi)\n"
    "    movq %rsi, %rsp\n"
    "    popq %r15\n"
    "    popq %r14\n"
    "    popq %r13\n"
    "    popq %r12\n"
    "    popq %rbx\n"
    "    popq %rbp\n"
    "    ret\n"
    ".size emu_switch,.-emu_switch\n");

double emu_now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

namespace emu {

Idx g_block_idx, g_block_dim, g_grid_dim;

namespace {
constexpr size_t STACK_BYTES = 256 << 10;
constexpr int MAX_THREADS = 1024;

struct Barrier {
    unsigned arrived = 0, generation = 0;
};
struct Fibre {
    void *sp = nullptr;
    bool done = true;
    const char *waiting = nullptr;       // what it is blocked on (diagnostics)
    Idx tid{ 0, 0, 0 };
};

Fibre g_fibres[MAX_THREADS];
char *g_stacks = nullptr;
void *g_sched_sp = nullptr;
int g_cur = -1, g_nthreads = 0, g_alive = 0;
unsigned long long g_progress = 0;       // bumped whenever a thread arrives at a barrier, passes one or exits
Barrier g_block_bar;
std::map<unsigned long long, Barrier> g_warp_bars;     // (warp << 32 | mask)
std::map<int, Barrier> g_named_bars;
uint64_t g_mail[MAX_THREADS];
alignas(128) unsigned char g_dyn_smem[232448];
void (*g_body)(void *) = nullptr;
void *g_arg = nullptr;

void yield_to_scheduler() { emu_switch(&g_fibres[g_cur].sp, g_sched_sp); }

void fibre_entry()
{
    g_body(g_arg);
    Fibre &f = g_fibres[g_cur];
    f.done = true;
    g_alive--;
    g_progress++;
    if (g_block_bar.arrived > 0 && g_block_bar.arrived >= (unsigned)g_alive) {     // the others wait in __syncthreads
        g_block_bar.arrived = 0;
        g_block_bar.generation++;
    }
    yield_to_scheduler();
    fprintf(stderr, "cuda_emu: a finished fibre was resumed\n");
    abort();
}

void wait_on(Barrier &b, unsigned expected, const char *what)
{
    g_progress++;
    const unsigned gen = b.generation;
    if (++b.arrived >= expected) {
        b.arrived = 0;
        b.generation++;
        return;
    }
    Fibre &f = g_fibres[g_cur];
    f.waiting = what;
    while (b.generation == gen) yield_to_scheduler();
    f.waiting = nullptr;
    g_progress++;
}

void report_deadlock()
{
    fprintf(stderr, "cuda_emu: DEADLOCK in block (%u,%u,%u) of a %ux%ux%u grid, %d of %d threads alive; on a GPU this kernel hangs\n",
            g_block_idx.x, g_block_idx.y, g_block_idx.z, g_grid_dim.x, g_grid_dim.y, g_grid_dim.z, g_alive, g_nthreads);
    int shown = 0;
    for (int t = 0; t < g_nthreads && shown < 48; t++)
        if (!g_fibres[t].done) {
            fprintf(stderr, "  thread %d waits at %s\n", t, g_fibres[t].waiting ? g_fibres[t].waiting : "?");
            shown++;
        }
    abort();
}
}   // namespace

Idx cur_tid() { return g_fibres[g_cur].tid; }
int cur_linear_tid() { return g_cur; }
uint64_t *warp_mail() { return g_mail + (g_cur & ~31); }
unsigned char *dyn_smem() { return g_dyn_smem; }

void sync_threads()
{
    // exited threads do not take part (Volta and later)
    wait_on(g_block_bar, (unsigned)g_alive, "__syncthreads");
}

void sync_warp(unsigned mask)
{
    const int warp = g_cur >> 5;
    const int lanes_in_block = std::min(32, g_nthreads - warp * 32);
    if (lanes_in_block < 32) mask &= (1u << lanes_in_block) - 1;
    if (!((mask >> (g_cur & 31)) & 1)) {
        fprintf(stderr, "cuda_emu: thread %d calls a warp primitive with mask %08x that does not name it\n", g_cur, mask);
        abort();
    }
    const unsigned expected = (unsigned)__builtin_popcount(mask);
    if (expected == 1) return;
    wait_on(g_warp_bars[((unsigned long long)warp << 32) | mask], expected, "a warp barrier / shuffle");
}

void sync_named(int id, int nthreads) { wait_on(g_named_bars[id], (unsigned)nthreads, "bar.sync (named barrier)"); }

void run_grid(dim3 grid, dim3 block, size_t smem, void (*body)(void *), void *arg)
{
    // EMU_ORDER=reverse schedules the threads of a block (and the blocks of a grid) from the last to the first: a
    // result that depends on the order in which threads reach a barrier-free stretch of code is a race on the GPU
    static const bool g_reverse = getenv("EMU_ORDER") && !strcmp(getenv("EMU_ORDER"), "reverse");
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > MAX_THREADS || smem > sizeof(g_dyn_smem)) {
        fprintf(stderr, "cuda_emu: invalid launch configuration (%d threads, %zu bytes of shared memory)\n", nthreads, smem);
        abort();
    }
    if (g_cur >= 0) {
        fprintf(stderr, "cuda_emu: kernel launch from device code is not emulated\n");
        abort();
    }
    if (!g_stacks) {
        g_stacks = (char *)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == MAP_FAILED) {
            perror("cuda_emu: mmap");
            abort();
        }
    }
    g_body = body;
    g_arg = arg;
    g_grid_dim = Idx{ grid.x, grid.y, grid.z };
    g_block_dim = Idx{ block.x, block.y, block.z };
    g_nthreads = nthreads;
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    for (unsigned long long bi = 0; bi < nblocks; bi++) {
                const unsigned long long b = g_reverse ? nblocks - 1 - bi : bi;
                const unsigned bx = (unsigned)(b % grid.x), by = (unsigned)((b / grid.x) % grid.y), bz = (unsigned)(b / ((unsigned long long)grid.x * grid.y));
                g_block_idx = Idx{ bx, by, bz };
                g_block_bar = Barrier();
                g_warp_bars.clear();
                g_named_bars.clear();
                g_alive = nthreads;
                for (int t = 0; t < nthreads; t++) {
                    Fibre &f = g_fibres[t];
                    f.done = false;
                    f.waiting = nullptr;
                    f.tid = Idx{ (unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y) };
                    // fresh stack: six callee-saved registers, then fibre_entry as the return address, then the
                    // slot a caller's return address would occupy (keeps the ABI's 16-byte alignment at entry)
                    uintptr_t top = ((uintptr_t)(g_stacks + (size_t)(t + 1) * STACK_BYTES)) & ~(uintptr_t)15;
                    void **sp = (void **)(top - 8);
                    *--sp = (void *)fibre_entry;
                    for (int r = 0; r < 6; r++) *--sp = nullptr;
                    f.sp = sp;
                }
                while (g_alive > 0) {
                    const unsigned long long before = g_progress;
                    for (int k = 0; k < nthreads; k++) {
                        const int t = g_reverse ? nthreads - 1 - k : k;
                        if (g_fibres[t].done) continue;
                        g_cur = t;
                        emu_switch(&g_sched_sp, g_fibres[t].sp);
                    }
                    g_cur = -1;
                    if (g_progress == before) report_deadlock();
                }
                g_cur = -1;
            }
}

}   // namespace emu
