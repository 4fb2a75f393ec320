/* emu_engine.cpp -- the CPU SIMT engine and the stand-in CUDA runtime of the emulator build (TEST INFRASTRUCTURE, see
 * tests/simt_emu/README.md and cuda_runtime.h in this directory).
 *
 * Execution model: a launch runs its CTAs one after another (in blockIdx order).  Inside a CTA every GPU thread is a
 * fiber with its own stack (a six-register context switch on x86-64, ucontext elsewhere); the scheduler runs the lanes of a warp one at a time until each has finished or
 * parked at a collective.  When every live lane of a warp is parked at a warp collective the results are computed and
 * the lanes continue; when every live lane of the CTA is parked at __syncthreads the barrier opens.  Lanes that have
 * returned from the kernel do not take part (as on the hardware).  Because a lane only stops at collectives, code between
 * two collectives is atomic with respect to the other lanes, which makes plain read-modify-write a correct atomicAdd, and
 * spinning on a flag is only safe if the flag was set by an EARLIER CTA (true for the ticketed onesweep look-back).
 * The arithmetic is the host compiler's IEEE f32 with -ffp-contract=off, i.e. the contract of include/lys_detmath.h.
 */
#include "cuda_runtime.h"
#include <sys/mman.h>
#if defined(__x86_64__) && !defined(LYS_EMU_UCONTEXT)
/* a context switch is six callee-saved registers and the stack pointer (System V x86-64); glibc's swapcontext would add a
 * sigprocmask system call per switch, and a pass makes millions of switches */
#define LYS_EMU_ASM_SWITCH 1
struct emu_ctx { void *sp; };
extern "C" void lys_emu_switch(emu_ctx *from, emu_ctx *to);
asm(".text\n.globl lys_emu_switch\n.type lys_emu_switch,@function\nlys_emu_switch:\n"
    "    pushq %rbp\n    pushq %rbx\n    pushq %r12\n    pushq %r13\n    pushq %r14\n    pushq %r15\n"
    "    movq %rsp, (%rdi)\n    movq (%rsi), %rsp\n"
    "    popq %r15\n    popq %r14\n    popq %r13\n    popq %r12\n    popq %rbx\n    popq %rbp\n    ret\n"
    ".size lys_emu_switch, .-lys_emu_switch\n");
#else
#include <ucontext.h>
typedef ucontext_t emu_ctx;
#endif
#include <chrono>
#include <cstdio>
#include <vector>

namespace emu {

uint3 g_block_idx = {0, 0, 0}, g_block_dim = {1, 1, 1}, g_grid_dim = {1, 1, 1};

namespace {
enum St { ST_RUN = 0, ST_WARP, ST_CTA, ST_DONE };
struct Lane {
    emu_ctx ctx;
    uint3 tid;
    int st, op;
    uint32_t a, b, out;
};
const int MAX_LANES = 1024;
const size_t STACK_BYTES = 512 * 1024;
Lane *g_lanes = nullptr;
char *g_stacks = nullptr;
Lane *g_cur = nullptr;
emu_ctx g_sched;
const std::function<void()> *g_body = nullptr;
std::vector<char> g_smem;
Stats g_stats = {0, 0, 0, 0, 0};
uint64_t g_ballots = 0, g_ballot_bits = 0;      /* votes / ballots and the lanes that voted true (LYS_EMU_TRACE) */
uint3 g_no_tid = {0, 0, 0};

[[noreturn]] void fatal(const char *msg) {
    fprintf(stderr, "simt_emu: %s (block %u, lane %d)\n", msg, g_block_idx.x, g_cur ? (int)(g_cur - g_lanes) : -1);
    abort();
}
#ifdef LYS_EMU_ASM_SWITCH
void to_sched(Lane *me) { lys_emu_switch(&me->ctx, &g_sched); }
void to_lane(Lane *l) { lys_emu_switch(&g_sched, &l->ctx); }
void trampoline() {
    (*g_body)();
    g_cur->st = ST_DONE;
    to_sched(g_cur);
    abort();                        /* a finished lane is never resumed */
}
void init_lane(Lane &L, char *stack) {
    /* first switch-in pops six registers and returns into trampoline() with the stack 8 mod 16, as after a call */
    void **top = (void **)(((uintptr_t)stack + STACK_BYTES) & ~(uintptr_t)15);
    *--top = nullptr;               /* trampoline's (unused) return address */
    *--top = (void *)trampoline;
    for (int k = 0; k < 6; k++) *--top = nullptr;
    L.ctx.sp = top;
}
#else
void to_sched(Lane *me) { swapcontext(&me->ctx, &g_sched); }
void to_lane(Lane *l) { swapcontext(&g_sched, &l->ctx); }
void trampoline() {
    (*g_body)();
    g_cur->st = ST_DONE;            /* returning resumes uc_link = the scheduler */
}
void init_lane(Lane &L, char *stack) {
    getcontext(&L.ctx);
    L.ctx.uc_stack.ss_sp = stack;
    L.ctx.uc_stack.ss_size = STACK_BYTES;
    L.ctx.uc_link = &g_sched;
    makecontext(&L.ctx, trampoline, 0);
}
#endif
void park(int st) {
    Lane *me = g_cur;
    if (!me) fatal("collective called outside a kernel");
    me->st = st;
    to_sched(me);
}
void resolve_warp(Lane *w, int n) {
    int op = 0;
    for (int l = 0; l < n; l++) if (w[l].st == ST_WARP) { if (!op) op = w[l].op; else if (op != w[l].op) fatal("lanes of one warp wait at different collectives"); }
    g_stats.warp_collectives++;
    switch (op) {
        case OP_BALLOT: {
            uint32_t m = 0;
            for (int l = 0; l < n; l++) if (w[l].st == ST_WARP && (w[l].a & 1u)) m |= 1u << l;
            for (int l = 0; l < n; l++) w[l].out = m;
            g_ballots++; g_ballot_bits += (uint64_t)__builtin_popcount(m);
            break;
        }
        case OP_SHFL: case OP_SHFL_XOR:
            for (int l = 0; l < n; l++) {
                if (w[l].st != ST_WARP) continue;
                int src = (op == OP_SHFL) ? (int)(w[l].b & 31u) : (l ^ (int)(w[l].b & 31u));
                w[l].out = (src < n && w[src].st == ST_WARP) ? w[src].a : w[l].a;
            }
            break;
        case OP_REDUCE_ADD: {
            uint32_t s = 0;
            for (int l = 0; l < n; l++) if (w[l].st == ST_WARP) s += w[l].a;
            for (int l = 0; l < n; l++) w[l].out = s;
            break;
        }
        case OP_REDUCE_MAX: {           /* signed maximum (__reduce_max_sync on int) */
            int32_t s = INT32_MIN;
            for (int l = 0; l < n; l++) if (w[l].st == ST_WARP && (int32_t)w[l].a > s) s = (int32_t)w[l].a;
            for (int l = 0; l < n; l++) w[l].out = (uint32_t)s;
            break;
        }
        case OP_MATCH_ANY:
            for (int l = 0; l < n; l++) {
                if (w[l].st != ST_WARP) continue;
                uint32_t m = 0;
                for (int k = 0; k < n; k++) if (w[k].st == ST_WARP && w[k].a == w[l].a) m |= 1u << k;
                w[l].out = m;
            }
            break;
        case OP_SYNCWARP: break;
        default: fatal("unknown collective");
    }
    for (int l = 0; l < n; l++) if (w[l].st == ST_WARP) w[l].st = ST_RUN;
}
/* LYS_EMU_SCHEDULE: 0 = CTAs, warps and lanes in index order (default); 1 = all three reversed; s >= 2 = pseudo-random orders
 * from seed s, redrawn for every CTA.  Results must not depend on it: a kernel whose output changes with the schedule has a
 * race (a missing barrier, an unordered read-modify-write) or an order-dependent result. */
int g_schedule = -1;
uint64_t g_rng = 0;
uint32_t next_rand() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (uint32_t)(g_rng >> 11); }
void make_order(int *ord, int n) {
    for (int i = 0; i < n; i++) ord[i] = (g_schedule == 1) ? n - 1 - i : i;
    if (g_schedule >= 2) for (int i = n - 1; i > 0; i--) { int j = (int)(next_rand() % (uint32_t)(i + 1)); int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
}
void run_cta(int n) {
    if (n > MAX_LANES) fatal("CTA larger than 1024 threads");
    if (!g_lanes) {
        g_lanes = new Lane[MAX_LANES];
        g_stacks = (char *)mmap(nullptr, STACK_BYTES * MAX_LANES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_stacks == (char *)MAP_FAILED) fatal("mmap of the lane stacks failed");
    }
    const unsigned bx = g_block_dim.x, by = g_block_dim.y;
    for (int t = 0; t < n; t++) {
        Lane &L = g_lanes[t];
        init_lane(L, g_stacks + (size_t)t * STACK_BYTES);
        L.tid.x = (unsigned)t % bx; L.tid.y = ((unsigned)t / bx) % by; L.tid.z = (unsigned)t / (bx * by);
        L.st = ST_RUN; L.op = 0; L.a = L.b = L.out = 0;
    }
    g_stats.ctas++; g_stats.lanes += (uint64_t)n;
    const int nw = (n + 31) / 32;
    int worder[32], lorder[32];
    make_order(worder, nw); make_order(lorder, 32);
    while (true) {
        bool progress = false;
        int done = 0, at_cta = 0;
        for (int wi = 0; wi < nw; wi++) {
            const int w = worder[wi];
            Lane *W = g_lanes + 32 * w;
            const int wn = (n - 32 * w < 32) ? n - 32 * w : 32;
            while (true) {
                for (int li = 0; li < 32; li++) {
                    const int l = lorder[li];
                    if (l < wn && W[l].st == ST_RUN) { g_cur = &W[l]; to_lane(&W[l]); g_cur = nullptr; progress = true; }
                }
                int nW = 0, nC = 0, nD = 0;
                for (int l = 0; l < wn; l++) { nW += W[l].st == ST_WARP; nC += W[l].st == ST_CTA; nD += W[l].st == ST_DONE; }
                if (nW == 0) { done += nD; at_cta += nC; break; }
                if (nC) fatal("a warp has lanes at a warp collective and lanes at __syncthreads");
                resolve_warp(W, wn);
            }
        }
        if (done == n) break;
        if (at_cta + done == n) {
            for (int t = 0; t < n; t++) if (g_lanes[t].st == ST_CTA) g_lanes[t].st = ST_RUN;
            g_stats.cta_barriers++;
            progress = true;
        }
        if (!progress) fatal("deadlock: no lane can run");
    }
}
}  // namespace

const uint3 &cur_tid() { return g_cur ? g_cur->tid : g_no_tid; }
uint32_t warp_collective(Op op, uint32_t a, uint32_t b) {
    Lane *me = g_cur;
    if (!me) fatal("warp collective called outside a kernel");
    me->op = op; me->a = a; me->b = b;
    park(ST_WARP);
    return me->out;
}
void cta_barrier() { park(ST_CTA); }
void *dyn_smem() { return g_smem.data(); }
Stats stats() { return g_stats; }

/* LYS_EMU_TRACE=1: one line per launch on stderr (kernel, grid, block, warp collectives, CTA barriers, mean number of lanes
 * that voted true) -- the number of collectives of a traversal kernel is its number of lock-step loop iterations, a proxy
 * for issued warp instructions, and its votes ask "is this lane still walking", so the last figure is its SIMT occupancy */
void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    static const bool trace = getenv("LYS_EMU_TRACE") && atoi(getenv("LYS_EMU_TRACE")) > 0;
    const Stats s0 = g_stats;
    const uint64_t b0 = g_ballots, bb0 = g_ballot_bits;
    if (g_cur) fatal("nested launch");
    const int n = (int)(block.x * block.y * block.z);
    if (n <= 0 || grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    g_smem.assign(smem + 16, 0);
    g_body = &body;
    g_grid_dim = {grid.x, grid.y, grid.z};
    g_block_dim = {block.x, block.y, block.z};
    g_stats.launches++;
    if (g_schedule < 0) { const char *e = getenv("LYS_EMU_SCHEDULE"); g_schedule = (e && atoi(e) > 0) ? atoi(e) : 0; g_rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)g_schedule; }
    const unsigned total = grid.x * grid.y * grid.z;
    std::vector<int> corder;
    if (g_schedule > 0 && total <= (1u << 24)) { corder.resize(total); make_order(corder.data(), (int)total); }
    for (unsigned c = 0; c < total; c++) {
        const unsigned k = corder.empty() ? c : (unsigned)corder[c];
        g_block_idx = {k % grid.x, (k / grid.x) % grid.y, k / (grid.x * grid.y)};
        run_cta(n);
    }
    g_body = nullptr;
    if (trace) fprintf(stderr, "emu launch %-28s grid %6u block %4d  warp collectives %10llu  cta barriers %8llu  true lanes per vote %5.1f\n", name,
                       grid.x * grid.y * grid.z, n, (unsigned long long)(g_stats.warp_collectives - s0.warp_collectives),
                       (unsigned long long)(g_stats.cta_barriers - s0.cta_barriers), g_ballots > b0 ? (double)(g_ballot_bits - bb0) / (double)(g_ballots - b0) : 0.0);
}

}  // namespace emu

/* ------------------------------------------------------------------ stand-in runtime */
struct emu_stream { int id; };
struct emu_event { double ms; };
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int env_int(const char *name, int dflt) { const char *e = getenv(name); return (e && *e) ? atoi(e) : dflt; }

const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"); }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = env_int("LYS_EMU_DEVICES", 1); return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceFlags(unsigned int *f) { *f = 0; return cudaSuccess; }
cudaError_t cudaSetDeviceFlags(unsigned int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof *p); strcpy(p->name, "lys SIMT emulator (CPU)"); p->multiProcessorCount = env_int("LYS_EMU_SMS", 4); return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = env_int("LYS_EMU_SMS", 4); return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { void *q = nullptr; if (posix_memalign(&q, 256, n ? n : 1)) { *p = nullptr; return cudaErrorMemoryAllocation; } *p = q; return cudaSuccess; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emu_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0.0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { if (e) e->ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }

extern "C" void lys_emu_stats(uint64_t *out5) {
    emu::Stats s = emu::stats();
    out5[0] = s.launches; out5[1] = s.ctas; out5[2] = s.lanes; out5[3] = s.warp_collectives; out5[4] = s.cta_barriers;
}
