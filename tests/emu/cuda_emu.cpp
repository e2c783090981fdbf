/*
 * cuda_emu.cpp — fiber scheduler behind cuda_emu.h.  TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
 */
#include "cuda_emu.h"

#include <execinfo.h>
#include <vector>

namespace emu {

[[noreturn]] static void die() {
    void *bt[32];
    int n = backtrace(bt, 32);
    backtrace_symbols_fd(bt, n, 2);
    std::abort();
}

ThreadCtx *cur = nullptr;
dim3 g_block, g_grid;
unsigned char *dyn_smem = nullptr;
static unsigned long long g_launches = 0;
static const char *g_kernel_name = "?";
unsigned long long launches() { return g_launches; }

/* ---- minimal x86-64 context switch (callee-saved registers + stack pointer) ------------------ */
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

struct Fiber {
    void *sp = nullptr;
    unsigned char *stack = nullptr;
    ThreadCtx ctx;
    bool done = false;
    const bool *wait_flag = nullptr; /* non-null: blocked until *wait_flag */
};

struct LaneSlot {
    int op = -1;
    unsigned mask = 0;
    unsigned long long value = 0;
    int arg = 0;
    unsigned long long result = 0;
    bool ready = false;
};

struct WarpState {
    unsigned arrived = 0;
    unsigned exited = 0;
    unsigned nlanes = 32;
    LaneSlot lane[32];
};

static const size_t kStack = 256 * 1024;
static std::vector<Fiber> g_fibers;
static std::vector<WarpState> g_warps;
static void *g_sched_sp = nullptr;
static Fiber *g_running = nullptr;
static const std::function<void()> *g_body = nullptr;
static unsigned g_barrier_arrived = 0, g_exited_threads = 0, g_nthreads = 0;
static bool g_barrier_release_flag[2] = {false, false};
static int g_barrier_phase = 0;

static void yield_to_scheduler() {
    Fiber *f = g_running;
    emu_switch(&f->sp, g_sched_sp);
}

static void fiber_main() {
    Fiber *f = g_running;
    (*g_body)();
    f->done = true;
    /* thread exit: release waiters that only needed the living threads */
    WarpState &w = g_warps[f->ctx.linear / 32];
    w.exited |= 1u << (f->ctx.linear % 32);
    ++g_exited_threads;
    if (g_barrier_arrived > 0 && g_barrier_arrived + g_exited_threads == g_nthreads) {
        g_barrier_release_flag[g_barrier_phase] = true;
        g_barrier_arrived = 0;
        g_barrier_phase ^= 1;
        g_barrier_release_flag[g_barrier_phase] = false;
    }
    if (w.arrived & w.exited) { /* cannot happen: a lane is either waiting or exited */
        std::fprintf(stderr, "cuda_emu: internal error\n");
        die();
    }
    /* a collective that names an exited lane can never complete: report it */
    for (unsigned l = 0; l < 32; ++l)
        if ((w.arrived >> l) & 1u)
            if (w.lane[l].mask & w.exited) {
                std::fprintf(stderr, "cuda_emu: warp collective (op %d, mask %08x) names lane(s) %08x that exited (block %u, warp %u)\n",
                             w.lane[l].op, w.lane[l].mask, w.lane[l].mask & w.exited, f->ctx.bid.x, f->ctx.linear / 32);
                die();
            }
    yield_to_scheduler();
    std::abort(); /* never resumed */
}

static void trampoline() { fiber_main(); }

static void make_fiber(Fiber &f) {
    if (!f.stack) f.stack = (unsigned char *)std::malloc(kStack);
    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;                 /* fake return address of trampoline */
    *--sp = (void *)&trampoline;     /* `ret` target */
    for (int i = 0; i < 6; ++i) *--sp = nullptr; /* rbp rbx r12 r13 r14 r15 */
    f.sp = (void *)sp;
    f.done = false;
    f.wait_flag = nullptr;
}

static unsigned long long finish_collective(WarpState &w, unsigned mask) {
    /* all lanes of `mask` have arrived: check agreement and compute per-lane results */
    int op = -1;
    int first = __builtin_ctz(mask);
    op = w.lane[first].op;
    for (unsigned l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) {
            if (w.lane[l].op != op || w.lane[l].mask != mask) {
                std::fprintf(stderr, "cuda_emu: mismatched warp collective: lane %u op %d mask %08x vs lane %d op %d mask %08x\n", l,
                             w.lane[l].op, w.lane[l].mask, first, op, mask);
                die();
            }
        }
    unsigned ballot = 0;
    if (op == OP_BALLOT || op == OP_ANY || op == OP_ALL)
        for (unsigned l = 0; l < 32; ++l)
            if (((mask >> l) & 1u) && w.lane[l].value) ballot |= 1u << l;
    for (unsigned l = 0; l < 32; ++l) {
        if (!((mask >> l) & 1u)) continue;
        LaneSlot &s = w.lane[l];
        int src = (int)l;
        {
            /* CUDA semantics with a segment width w (power of two): lanes are grouped in segments of
             * w; a source outside the caller's segment returns the caller's own value (up/down/xor)
             * or wraps inside the segment (indexed) */
            const int operand = s.arg & 0xFFFF;
            int w = (s.arg >> 16) & 0xFFFF;
            if (w <= 0 || w > 32) w = 32;
            const int seg = ((int)l / w) * w;
            switch (op) {
            case OP_SHFL: src = seg + (operand % w); break;
            case OP_SHFL_UP: src = (int)l - operand; if (src < seg) src = (int)l; break;
            case OP_SHFL_DOWN: src = (int)l + operand; if (src >= seg + w) src = (int)l; break;
            case OP_SHFL_XOR: src = (int)l ^ operand; if (src < seg || src >= seg + w) src = (int)l; break;
            default: break;
            }
        }
        switch (op) {
        case OP_SHFL: case OP_SHFL_UP: case OP_SHFL_DOWN: case OP_SHFL_XOR:
            if (!((mask >> src) & 1u)) {
                /* reading from a lane outside the mask is undefined in CUDA */
                if (op == OP_SHFL) {
                    std::fprintf(stderr, "cuda_emu: __shfl_sync reads lane %d which is not in mask %08x\n", src, mask);
                    die();
                }
                src = (int)l;
            }
            s.result = w.lane[src].value;
            break;
        case OP_BALLOT: s.result = ballot; break;
        case OP_ANY: s.result = ballot != 0; break;
        case OP_ALL: s.result = ballot == mask; break;
        default: s.result = 0; break;
        }
    }
    for (unsigned l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) w.lane[l].ready = true;
    w.arrived &= ~mask;
    return 0;
}

unsigned long long warp_collective(int op, unsigned mask, unsigned long long value, int arg) {
    Fiber *f = g_running;
    const unsigned lane = f->ctx.linear % 32;
    WarpState &w = g_warps[f->ctx.linear / 32];
    const unsigned present = (w.nlanes >= 32) ? 0xffffffffu : ((1u << w.nlanes) - 1u);
    mask &= present; /* partial last warp: CUDA ignores non-existent lanes */
    if (!((mask >> lane) & 1u)) {
        std::fprintf(stderr, "cuda_emu: lane %u calls a collective with mask %08x that excludes it\n", lane, mask);
        die();
    }
    if (mask & w.exited) {
        std::fprintf(stderr, "cuda_emu: [%s] collective op %d mask %08x includes exited lanes %08x (block %u thread %u)\n", g_kernel_name, op, mask, mask & w.exited, f->ctx.bid.x, f->ctx.linear);
        die();
    }
    LaneSlot &s = w.lane[lane];
    s.op = op; s.mask = mask; s.value = value; s.arg = arg; s.ready = false;
    w.arrived |= 1u << lane;
    if ((w.arrived & mask) == mask) finish_collective(w, mask);
    while (!s.ready) {
        f->wait_flag = &s.ready;
        yield_to_scheduler();
    }
    f->wait_flag = nullptr;
    s.ready = false;
    return s.result;
}

void block_barrier() {
    Fiber *f = g_running;
    const int phase = g_barrier_phase;
    ++g_barrier_arrived;
    if (g_barrier_arrived + g_exited_threads == g_nthreads) {
        g_barrier_release_flag[phase] = true;
        g_barrier_arrived = 0;
        g_barrier_phase ^= 1;
        g_barrier_release_flag[g_barrier_phase] = false;
        return;
    }
    while (!g_barrier_release_flag[phase]) {
        f->wait_flag = &g_barrier_release_flag[phase];
        yield_to_scheduler();
    }
    f->wait_flag = nullptr;
}

void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    ++g_launches;
    g_kernel_name = name;
    if (std::getenv("DSIM_EMU_TRACE")) std::fprintf(stderr, "cuda_emu: launch %s grid %u block %u\n", name, grid.x, block.x);
    g_grid = grid;
    g_block = block;
    g_body = &body;
    const unsigned nthreads = block.x * block.y * block.z;
    g_nthreads = nthreads;
    if (g_fibers.size() < nthreads) g_fibers.resize(nthreads);
    static std::vector<unsigned char> smem_buf;
    if (smem_buf.size() < smem + 16) smem_buf.resize(smem + 16);
    dyn_smem = (unsigned char *)(((uintptr_t)smem_buf.data() + 15) & ~(uintptr_t)15);
    const unsigned nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_warps.assign(nwarps, WarpState());
                for (unsigned wi = 0; wi < nwarps; ++wi) g_warps[wi].nlanes = std::min(32u, nthreads - wi * 32);
                g_barrier_arrived = 0; g_exited_threads = 0; g_barrier_phase = 0;
                g_barrier_release_flag[0] = g_barrier_release_flag[1] = false;
                for (unsigned t = 0; t < nthreads; ++t) {
                    Fiber &f = g_fibers[t];
                    make_fiber(f);
                    f.ctx.linear = t;
                    f.ctx.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    f.ctx.bid = uint3{bx, by, bz};
                }
                unsigned alive = nthreads;
                while (alive > 0) {
                    bool progress = false;
                    for (unsigned t = 0; t < nthreads; ++t) {
                        Fiber &f = g_fibers[t];
                        if (f.done) continue;
                        if (f.wait_flag && !*f.wait_flag) continue;
                        g_running = &f;
                        cur = &f.ctx;
                        emu_switch(&g_sched_sp, f.sp);
                        progress = true;
                        if (f.done) --alive;
                    }
                    if (!progress) {
                        std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %u threads blocked\n", bx, by, bz, alive);
                        for (unsigned wi = 0; wi < nwarps; ++wi)
                            std::fprintf(stderr, "  warp %u arrived %08x exited %08x\n", wi, g_warps[wi].arrived, g_warps[wi].exited);
                        die();
                    }
                }
            }
    g_running = nullptr;
    cur = nullptr;
}

} // namespace emu
