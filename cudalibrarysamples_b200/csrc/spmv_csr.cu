// spmv_csr.cu -- CSR  y = alpha*A*x + beta*y  for B200 (sm_100a), fp32 / fp64, int32 indices.
//
// Replaces the closed cusparse::partition_kernel / csrmv_v3_kernel / spmv_fixup_kernel trio behind
// cusparseSpMV_preprocess / cusparseSpMV (reference call sites: cuSPARSE/spmv_csr/spmv_csr_example.c:104-112,
// cuSPARSE/cg/cg_example.c:156,220,294,415, cuSPARSE/bicgstab/bicgstab_example.c:190,262,315,358,504).
//
// Design (see DESIGN.md "CSR kernel"):
//   * analyze: merge-path style partition of the (rows + nnz) item list into tiles of TILE_ITEMS items.
//     A tile boundary that falls inside a row shorter than LONG_ROW is rounded down to that row's start,
//     so ordinary rows never straddle tiles and need no carry/fix-up; only rows >= LONG_ROW are cut, and
//     their per-tile partial sums are stored and combined, in fixed tile order, by whichever CTA finishes last
//     (bit-reproducible, single launch, no spin-waits, nothing on the tiles' critical path).
//   * mv: one CTA per tile.  Phase 1 streams val[]/col_ind[] with L1-bypassing loads (every warp-level
//     load is one aligned 128 B / 256 B segment, several steps in flight before the first use), gathers x
//     through L1/L2 and parks the products in shared memory.  Phase 2 reduces each row from shared memory with a group of
//     g = 1..32 lanes (g picked per tile from its mean row length) and a warp-shuffle tree, and writes y
//     once.  No tensor cores: 0.125-0.17 flop/B, HBM-bound.
#include "spmv_common.cuh"
#include "../../include/b200spmv.h"

namespace b200 {

// Tunables (overridable with -D for the parameter sweeps in scripts/sweep.py)
#ifndef B200_CSR_TILE_ITEMS
#define B200_CSR_TILE_ITEMS 2048
#endif
#ifndef B200_CSR_LONG_ROW
#define B200_CSR_LONG_ROW 512
#endif
#ifndef B200_CSR_BLOCK
#define B200_CSR_BLOCK 256
#endif
#ifndef B200_CSR_BATCH
#define B200_CSR_BATCH 4
#endif
#ifndef B200_CSR_MIN_CTAS
#define B200_CSR_MIN_CTAS 4
#endif
#ifndef B200_CSR_KERNEL      // 0 = one CTA per tile, 1 = persistent software-pipelined CTAs
#define B200_CSR_KERNEL 1
#endif
#ifndef B200_CSR_PIPE_STEPS
#define B200_CSR_PIPE_STEPS ((B200_CSR_TILE_ITEMS + B200_CSR_BLOCK - 1) / B200_CSR_BLOCK)
#endif
#ifndef B200_CSR_PIPE_OFFS
#define B200_CSR_PIPE_OFFS 4
#endif
#ifndef B200_CSR_RED_ROWS     // rows a lane group reduces concurrently in phase 2
#define B200_CSR_RED_ROWS 4
#endif
#ifndef B200_CSR_RED_U        // predicated product loads per lane and row before the fall-back loop
#define B200_CSR_RED_U 2
#endif
#ifndef B200_CSR_ABLATE   // profiling only: 1 = skip phase 2, 2 = no x gather, 3 = neither (stream only)
#define B200_CSR_ABLATE 0
#endif
constexpr int CSR_TILE_ITEMS = B200_CSR_TILE_ITEMS;  // merge items (row ends + non-zeros) per tile
constexpr int CSR_LONG_ROW   = B200_CSR_LONG_ROW;    // rows at least this long may be split between tiles
constexpr int CSR_BLOCK      = B200_CSR_BLOCK;       // threads per CTA
constexpr int CSR_SMEM_ELEMS = CSR_TILE_ITEMS + CSR_LONG_ROW;  // max non-zeros a tile can hold
constexpr int CSR_BATCH      = B200_CSR_BATCH;  // load steps whose loads are all issued before the first use
// a tile's element range starts at a multiple of 32 (<= 31 masked lanes) and holds < CSR_SMEM_ELEMS non-zeros
constexpr int CSR_STEPS      = (CSR_SMEM_ELEMS + 31 + CSR_BLOCK - 1) / CSR_BLOCK;
constexpr int CSR_ITERS      = (CSR_STEPS + CSR_BATCH - 1) / CSR_BATCH * CSR_BATCH;

constexpr size_t PLAN_HEADER_BYTES = 256;

struct PlanView {
    int2*   tiles;      // [num_tiles+1] (row, nnz) start coordinate of each tile
    int*    ctl;        // [0] = CTAs that finished the current launch, [1] = number of split rows (set by analyze)
    int4*   split;      // [num_tiles+1] split rows: (row, first covering tile b1, last covering tile b2, -)
    double* head_part;  // [num_tiles+1] partial sum of the split row a tile starts in
    double* tail_part;  // [num_tiles+1] partial sum of the split row a tile ends in
};

static inline int64_t csr_num_tiles(int64_t rows, int64_t nnz) {
    return (rows + nnz + CSR_TILE_ITEMS - 1) / CSR_TILE_ITEMS;
}
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline size_t plan_layout(int64_t num_tiles, void* ws, PlanView* v) {
    size_t n = (size_t)num_tiles + 1;
    size_t o_tiles = PLAN_HEADER_BYTES;
    size_t o_ctl   = align_up(o_tiles + n * sizeof(int2), 256);
    size_t o_split = align_up(o_ctl + 64, 256);
    size_t o_head  = align_up(o_split + n * sizeof(int4), 256);
    size_t o_tail  = align_up(o_head + n * sizeof(double), 256);
    size_t total   = align_up(o_tail + n * sizeof(double), 256);
    if (v) {
        char* b = (char*)ws;
        v->tiles = (int2*)(b + o_tiles);
        v->ctl = (int*)(b + o_ctl);
        v->split = (int4*)(b + o_split);
        v->head_part = (double*)(b + o_head);
        v->tail_part = (double*)(b + o_tail);
    }
    return total;
}

// ------------------------------------------------------------------------------------------------
// analyze: one thread per tile boundary.
//   boundary b sits on merge diagonal d = min(b*TILE, rows+nnz); r = the row the diagonal cuts
//   (largest r with r + off[r] <= d).  Rows shorter than LONG_ROW are never cut: the boundary moves to
//   the row start (r, off[r]); long rows are cut exactly at the diagonal (r, off[r] + e).
// ------------------------------------------------------------------------------------------------
__global__ void csr_partition_kernel(const int* __restrict__ off, int base, int64_t rows, int64_t nnz,
                                     int64_t num_tiles, PlanView plan) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > num_tiles) return;
    int64_t d = b * CSR_TILE_ITEMS;
    if (d > rows + nnz) d = rows + nnz;
    int64_t lo = 0, hi = rows;
    while (lo < hi) {
        int64_t mid = (lo + hi + 1) >> 1;
        if (mid + (int64_t)(off[mid] - base) <= d) lo = mid; else hi = mid - 1;
    }
    int64_t r = lo, n = off[r] - base;
    if (r < rows) {
        int64_t len = (int64_t)(off[r + 1] - base) - n;
        int64_t e = d - (r + n);
        if (len >= CSR_LONG_ROW && e > 0) {
            // Row r is cut here.  It is covered by tiles b1 = g(r)/TILE (holding its start as a "tail") through
            // b2 = (g(r)+len)/TILE (the others hold "heads"); the boundary that makes the FIRST cut registers the row
            // in the split list that the fix-up pass walks.
            const int64_t g0 = r + n, b1 = g0 / CSR_TILE_ITEMS, b2 = (g0 + len) / CSR_TILE_ITEMS;
            if (b == b1 + 1) {
                const int slot = atomicAdd(plan.ctl + 1, 1);
                plan.split[slot] = make_int4((int)r, (int)b1, (int)b2, 0);
            }
            n += e;
        }
    }
    plan.tiles[b] = make_int2((int)r, (int)n);
    plan.head_part[b] = 0.0;
    plan.tail_part[b] = 0.0;
}

#ifdef B200_CSR_TRACE   // debugging builds only: per-tile clock64 stamps {start, after phase 1, end, smid}
#define TRACE_STAMP(a, b, slot) do { if ((a).trace && threadIdx.x == 0) (a).trace[(size_t)(b) * 4 + (slot)] = clock64(); } while (0)
#define TRACE_SMID(a, b) do { if ((a).trace && threadIdx.x == 0) { unsigned s_; asm volatile("mov.u32 %0, %%smid;" : "=r"(s_)); (a).trace[(size_t)(b) * 4 + 3] = s_; } } while (0)
static long long* g_trace_ptr = nullptr;
extern "C" void b200spmv_debug_set_trace(void* p) { g_trace_ptr = (long long*)p; }
#else
#define TRACE_STAMP(a, b, slot) do { } while (0)
#define TRACE_SMID(a, b) do { } while (0)
#endif

template <typename T>
struct CsrArgs {
    const int* off;
    const int* col;
    const T*   val;
    const T*   x;
    T*         y;
    int        base;
    int        rows;
    int        nnz;
    Scalars<T> s;
    PlanView   plan;
    long long* trace;
};

// Sum sP[lo, hi) with the whole CTA, fixed order (bit-reproducible). Result valid on thread 0.
template <typename T, int BLOCK>
__device__ __forceinline__ T block_sum_range(const T* sP, int lo, int hi, T* sRed) {
    T s = T(0);
    for (int k = lo + (int)threadIdx.x; k < hi; k += BLOCK) s += sP[k];
    s = warp_sum(s);
    __syncthreads();  // sRed reuse
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = s;
    __syncthreads();
    T tot = T(0);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < BLOCK / 32; w++) tot += sRed[w];
    }
    return tot;
}

// Rows [r_first, r_first + nrows) are complete inside the tile; the products of row r live at
// sP[sOff[r - rs] .. sOff[r - rs + 1]) (sOff = the tile's slice of rowOff, staged in phase 1, rebased to the tile).
//
// Every shared-memory access and every shuffle of this phase queues in the SM's single L1TEX FIFO behind the x
// gathers of the other resident CTAs, so what matters is the NUMBER OF DEPENDENT STEPS, not the instruction count:
// a group of G lanes therefore works on RED_ROWS rows at once -- all row bounds first, then all products (RED_U
// predicated loads per lane and row, enough for rows up to RED_U*G long), then RED_ROWS interleaved shuffle trees.
constexpr int RED_ROWS = B200_CSR_RED_ROWS;
constexpr int RED_U    = B200_CSR_RED_U;

template <typename T, int G, int ROWS, int U, int BLOCK>
__device__ __forceinline__ void reduce_rows(const CsrArgs<T>& a, const T* sP, const int* sOff, int rs, int r_first,
                                            int nrows, T alpha, T beta) {
    constexpr int GROUPS = BLOCK / G;
    const int gid = threadIdx.x / G, gl = threadIdx.x % G;
    const int* so = sOff + (r_first - rs);
    for (int r0 = 0; r0 < nrows; r0 += GROUPS * ROWS) {
        int k0[ROWS], e[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; i++) {
            const int ri = r0 + i * GROUPS + gid;
            const bool active = ri < nrows;
            k0[i] = (active ? so[ri] : 0) + gl;
            e[i]  = active ? so[ri + 1] : 0;
        }
        T p[ROWS][U];
#pragma unroll
        for (int i = 0; i < ROWS; i++)
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int k = k0[i] + u * G;
                p[i][u] = k < e[i] ? sP[k] : T(0);
            }
        T sum[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; i++) {
            sum[i] = p[i][0];
#pragma unroll
            for (int u = 1; u < U; u++) sum[i] += p[i][u];
        }
#pragma unroll
        for (int i = 0; i < ROWS; i++) {                        // rows longer than U*G: batches of U loads in flight
            for (int kb = k0[i] + U * G; kb < e[i]; kb += U * G) {
                T q[U];
#pragma unroll
                for (int u = 0; u < U; u++) q[u] = kb + u * G < e[i] ? sP[kb + u * G] : T(0);
#pragma unroll
                for (int u = 0; u < U; u++) sum[i] += q[u];
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < ROWS; i++) sum[i] += __shfl_down_sync(0xffffffffu, sum[i], o, G);
        if (gl == 0) {
#pragma unroll
            for (int i = 0; i < ROWS; i++) {
                const int ri = r0 + i * GROUPS + gid;
                if (ri < nrows) {
                    T* yp = a.y + r_first + ri;
                    *yp = axpby(alpha, sum[i], beta, yp);
                }
            }
        }
    }
}

// Split rows (>= LONG_ROW non-zeros, cut by a tile boundary): every covering tile only STORES its partial sum
// (head_part[b] / tail_part[b], fire-and-forget, nothing waits).  When a CTA has finished all its tiles it bumps
// ctl[0]; the CTA that arrives last walks the split list built by analyze and writes
//     y[R] = alpha * (tail_part[b1] + head_part[b1+1] + ... + head_part[b2]) + beta * y[R]
// in fixed tile order -> bit-reproducible, single launch, no spin-waits, counter left at 0 for the next launch.
template <typename T>
__device__ __forceinline__ void split_rows_fixup(const CsrArgs<T>& a, T alpha, T beta) {
    __shared__ int s_last;
    __threadfence();                       // this CTA's partial sums are visible device-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(a.plan.ctl, 1);
        s_last = old == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nsplit = __ldcg(a.plan.ctl + 1);
    for (int i = (int)threadIdx.x; i < nsplit; i += (int)blockDim.x) {
        const int4 sr = a.plan.split[i];
        double sum = __ldcg(a.plan.tail_part + sr.y);
        for (int b = sr.y + 1; b <= sr.z; b++) sum += __ldcg(a.plan.head_part + b);
        T* yp = a.y + sr.x;
        *yp = axpby(alpha, (T)sum, beta, yp);
    }
    if (threadIdx.x == 0) a.plan.ctl[0] = 0;
}

// ---------------- phase 2: per-row reduction out of shared memory ----------------------------------
// sP[0 .. ne-ns) holds the tile's products, sOff[i] = rowOff[rs+i] - base - ns for i = 0 .. re-rs.
template <typename T>
__device__ __forceinline__ void tile_phase2(const CsrArgs<T>& a, int b, int rs, int ns, int re, int ne, const T* sP,
                                            const int* sOff, T* sRed, T alpha, T beta) {
    if (B200_CSR_ABLATE & 1) { if (sP[threadIdx.x] == T(1.2345)) a.y[0] = sP[0]; return; }
    const int cnt = ne - ns;
    bool head = false;
    int  head_end = 0;  // products [0, head_end) belong to the split row rs
    if (rs < a.rows && sOff[0] < 0) {            // the tile starts inside row rs
        head = true;
        const int o1 = re > rs ? sOff[1] : cnt;  // re == rs: the whole tile lies inside row rs
        head_end = o1 < cnt ? o1 : cnt;
    }
    const int r_first = rs + (head ? 1 : 0);
    const int nrows   = re - r_first;  // complete rows (may be <= 0)
    int  tail_beg = cnt;               // products [tail_beg, cnt) belong to the split row re
    bool tail = false;
    if (re < a.rows && re >= r_first) {
        const int o0 = sOff[re - rs];
        if (cnt > o0) { tail = true; tail_beg = o0; }
    }

    if (nrows > 0) {
        const int body = tail_beg - head_end;
        const int avg2 = body / (2 * nrows);  // half the mean row length
        if      (avg2 <= 1)  reduce_rows<T, 1,  RED_ROWS, RED_U, CSR_BLOCK>(a, sP, sOff, rs, r_first, nrows, alpha, beta);
        else if (avg2 <= 2)  reduce_rows<T, 2,  RED_ROWS, RED_U, CSR_BLOCK>(a, sP, sOff, rs, r_first, nrows, alpha, beta);
        else if (avg2 <= 4)  reduce_rows<T, 4,  RED_ROWS, RED_U, CSR_BLOCK>(a, sP, sOff, rs, r_first, nrows, alpha, beta);
        else if (avg2 <= 8)  reduce_rows<T, 8,  RED_ROWS, RED_U, CSR_BLOCK>(a, sP, sOff, rs, r_first, nrows, alpha, beta);
        else if (avg2 <= 16) reduce_rows<T, 16, RED_ROWS, RED_U, CSR_BLOCK>(a, sP, sOff, rs, r_first, nrows, alpha, beta);
        else                 reduce_rows<T, 32, 2, 8, CSR_BLOCK>(a, sP, sOff, rs, r_first, nrows, alpha, beta);  // long rows: 256 elements per batch
    }

    if (head) {  // block-uniform
        const T hs = block_sum_range<T, CSR_BLOCK>(sP, 0, head_end, sRed);
        if (threadIdx.x == 0) a.plan.head_part[b] = (double)hs;
    }
    if (tail) {  // block-uniform
        const T ts = block_sum_range<T, CSR_BLOCK>(sP, tail_beg, cnt, sRed);
        if (threadIdx.x == 0) a.plan.tail_part[b] = (double)ts;
    }
}

template <typename T>
__global__ void __launch_bounds__(CSR_BLOCK, B200_CSR_MIN_CTAS) csr_tile_kernel(const CsrArgs<T> a) {
    __shared__ T   sP[CSR_SMEM_ELEMS];
    __shared__ int sOff[CSR_SMEM_ELEMS + 1];   // rowOff[rs .. re] - base - ns: a tile spans < CSR_SMEM_ELEMS rows
    __shared__ T   sRed[CSR_BLOCK / 32];

    const int  b  = blockIdx.x;
    const int2 st = a.plan.tiles[b], en = a.plan.tiles[b + 1];
    const int  rs = st.x, ns = st.y, re = en.x, ne = en.y;
    const T alpha = a.s.a(), beta = a.s.b();
    TRACE_STAMP(a, b, 0); TRACE_SMID(a, b);

    // ---------------- phase 1: stream val/col, gather x, park products in shared memory -------------
    // Lane l of a warp handles element (step*BLOCK + warp*32 + l): every load/gather instruction covers 32
    // CONSECUTIVE non-zeros.  The streaming loads are then perfectly coalesced (the tile's first element is
    // rounded down to a multiple of 32 so each warp-level load is one aligned 128 B / 256 B segment), the
    // gathers of x see the sorted neighbouring columns of a row in one instruction (fewest L1TEX wavefronts:
    // the gather rate of the SM, ~1 distinct 128 B line per clock, is what bounds this kernel on scattered
    // matrices), and the shared-memory stores are bank-conflict free.
    {
        // the tile's slice of rowOff goes to shared memory first: its latency overlaps the val/col stream, and
        // phase 2 then never waits on global memory
        const int noff = (re < a.rows ? re : a.rows) - rs + 1;
        for (int i = (int)threadIdx.x; i < noff; i += CSR_BLOCK) sOff[i] = __ldg(a.off + rs + i) - a.base - ns;
        const int al   = ns & ~31;
        const int lead = ns - al;                 // masked lanes in front of the tile
        const int span = ne - al;                 // elements [lead, span) are live
        const int* colp = a.col + al;
        const T*   valp = a.val + al;
#pragma unroll
        for (int batch = 0; batch < CSR_ITERS; batch += CSR_BATCH) {
            if (batch * CSR_BLOCK < span) {       // block-uniform
                int c[CSR_BATCH];
                T   v[CSR_BATCH], xv[CSR_BATCH];
#pragma unroll
                for (int k = 0; k < CSR_BATCH; k++) {
                    const int e = (batch + k) * CSR_BLOCK + (int)threadIdx.x;
                    const bool live = e >= lead && e < span;
                    c[k] = live ? ldg_stream(colp + e) : a.base;
                    v[k] = live ? ldg_stream(valp + e) : T(0);
                }
#pragma unroll
                for (int k = 0; k < CSR_BATCH; k++) {
                    const int e = (batch + k) * CSR_BLOCK + (int)threadIdx.x;
                    const bool live = e >= lead && e < span;
                    xv[k] = (live && !(B200_CSR_ABLATE & 2)) ? __ldg(a.x + (c[k] - a.base)) : T(c[k]);
                }
#pragma unroll
                for (int k = 0; k < CSR_BATCH; k++) {
                    const int e = (batch + k) * CSR_BLOCK + (int)threadIdx.x;
                    if (e >= lead && e < span) sP[e - lead] = v[k] * xv[k];
                }
            }
        }
    }
    __syncthreads();
    TRACE_STAMP(a, b, 1);

    tile_phase2<T>(a, b, rs, ns, re, ne, sP, sOff, sRed, alpha, beta);
    TRACE_STAMP(a, b, 2);
    split_rows_fixup<T>(a, alpha, beta);
}

// ================================================================================================
// Persistent, software-pipelined variant (the default): each CTA walks tiles b, b+grid, b+2*grid, ...
// While it reduces tile i out of shared memory (phase 2), the val/col/rowOff loads of tile i+1 are
// already in flight into registers, so the HBM latency of the stream never sits on a tile's critical
// path; what remains per tile is one L2 round trip for the x gathers plus the shared-memory work.
// No CTA launch gaps either: the grid is SMs x (resident CTAs per SM).
// ================================================================================================
constexpr int PIPE_STEPS = B200_CSR_PIPE_STEPS;   // load steps held in registers for the next tile
constexpr int PIPE_OFFS  = B200_CSR_PIPE_OFFS;    // rowOff entries per thread held in registers

template <typename T>
struct TileRegs {
    int c[PIPE_STEPS];
    T   v[PIPE_STEPS];
    int o[PIPE_OFFS];
};

template <typename T>
__device__ __forceinline__ void pipe_issue_loads(const CsrArgs<T>& a, int2 st, int2 en, TileRegs<T>& t) {
    const int rs = st.x, ns = st.y, re = en.x, ne = en.y;
    const int noff = re - rs + 1;
#pragma unroll
    for (int j = 0; j < PIPE_OFFS; j++) {
        const int i = j * CSR_BLOCK + (int)threadIdx.x;
        t.o[j] = i < noff ? __ldg(a.off + rs + i) : 0;
    }
    const int al = ns & ~31, lead = ns - al, span = ne - al;
    const int* colp = a.col + al;
    const T*   valp = a.val + al;
#pragma unroll
    for (int k = 0; k < PIPE_STEPS; k++) {
        const int e = k * CSR_BLOCK + (int)threadIdx.x;
        const bool live = e >= lead && e < span;
        t.c[k] = live ? ldg_stream(colp + e) : a.base;
        t.v[k] = live ? ldg_stream(valp + e) : T(0);
    }
}

template <typename T>
__global__ void __launch_bounds__(CSR_BLOCK, B200_CSR_MIN_CTAS) csr_pipe_kernel(const CsrArgs<T> a, int num_tiles) {
    __shared__ T   sP[CSR_SMEM_ELEMS];
    __shared__ int sOff[CSR_SMEM_ELEMS + 1];
    __shared__ T   sRed[CSR_BLOCK / 32];

    int b = blockIdx.x;                      // the launcher guarantees gridDim.x <= num_tiles
    const T alpha = a.s.a(), beta = a.s.b();
    int2 st = a.plan.tiles[b], en = a.plan.tiles[b + 1];
    TileRegs<T> t;
    pipe_issue_loads(a, st, en, t);

    for (;;) {
        const int rs = st.x, ns = st.y, re = en.x, ne = en.y;
        TRACE_STAMP(a, b, 0); TRACE_SMID(a, b);
        const int nb = b + (int)gridDim.x;
        const bool has_next = nb < num_tiles;
        int2 nst = st, nen = en;
        if (has_next) { nst = a.plan.tiles[nb]; nen = a.plan.tiles[nb + 1]; }   // descriptor of the next tile, early

        // rowOff slice -> shared memory (rebased to the tile)
        const int noff = re - rs + 1;
#pragma unroll
        for (int j = 0; j < PIPE_OFFS; j++) {
            const int i = j * CSR_BLOCK + (int)threadIdx.x;
            if (i < noff) sOff[i] = t.o[j] - a.base - ns;
        }
        for (int i = PIPE_OFFS * CSR_BLOCK + (int)threadIdx.x; i < noff; i += CSR_BLOCK)      // very row-dense tiles only
            sOff[i] = __ldg(a.off + rs + i) - a.base - ns;

        // gathers + products
        const int al = ns & ~31, lead = ns - al, span = ne - al;
        {
            T xv[PIPE_STEPS];
#pragma unroll
            for (int k = 0; k < PIPE_STEPS; k++) {
                const int e = k * CSR_BLOCK + (int)threadIdx.x;
                const bool live = e >= lead && e < span;
                xv[k] = (live && !(B200_CSR_ABLATE & 2)) ? __ldg(a.x + (t.c[k] - a.base)) : T(0);
            }
#pragma unroll
            for (int k = 0; k < PIPE_STEPS; k++) {
                const int e = k * CSR_BLOCK + (int)threadIdx.x;
                if (e >= lead && e < span) sP[e - lead] = t.v[k] * xv[k];
            }
        }
        for (int e = PIPE_STEPS * CSR_BLOCK + (int)threadIdx.x; e < span; e += CSR_BLOCK) {   // oversized tiles only
            const int c = ldg_stream(a.col + al + e);
            const T   v = ldg_stream(a.val + al + e);
            sP[e - lead] = v * __ldg(a.x + (c - a.base));
        }
        __syncthreads();
        TRACE_STAMP(a, b, 1);

        if (has_next) pipe_issue_loads(a, nst, nen, t);   // block-uniform; overlaps phase 2 below

        tile_phase2<T>(a, b, rs, ns, re, ne, sP, sOff, sRed, alpha, beta);
        TRACE_STAMP(a, b, 2);
        if (!has_next) break;
        __syncthreads();                                   // sP / sOff are free again
        b = nb; st = nst; en = nen;
    }
    split_rows_fixup<T>(a, alpha, beta);
}

// SMs x resident CTAs per SM of a kernel on the current device (cached per device / kernel).
static int resident_ctas(const void* kernel) {
    struct Entry { int dev; const void* k; int n; };
    static Entry cache[32];
    static int ncache = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < ncache; i++)
        if (cache[i].dev == dev && cache[i].k == kernel) return cache[i].n;
    int sms = 148, per = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, CSR_BLOCK, 0);
    if (per < 1) per = 1;
    const int n = sms * per;
    if (ncache < 32) cache[ncache++] = Entry{dev, kernel, n};
    return n;
}

template <typename T>
static int launch_csr(cudaStream_t stream, int64_t rows, int64_t nnz, const void* off, const void* col,
                      const void* val, int base, const void* alpha, const void* beta, int on_device,
                      const void* x, void* y, void* ws) {
    const int64_t nt = csr_num_tiles(rows, nnz);
    if (nt == 0) return 0;
    CsrArgs<T> a;
    a.off = (const int*)off; a.col = (const int*)col; a.val = (const T*)val;
    a.x = (const T*)x; a.y = (T*)y; a.base = base; a.rows = (int)rows; a.nnz = (int)nnz;
    if (on_device) { a.s.alpha = T(0); a.s.beta = T(0); a.s.alpha_dev = (const T*)alpha; a.s.beta_dev = (const T*)beta; }
    else { a.s.alpha = *(const T*)alpha; a.s.beta = *(const T*)beta; a.s.alpha_dev = nullptr; a.s.beta_dev = nullptr; }
    plan_layout(nt, ws, &a.plan);
    a.trace = nullptr;
#ifdef B200_CSR_TRACE
    a.trace = g_trace_ptr;
#endif
#if B200_CSR_KERNEL == 0
    csr_tile_kernel<T><<<(unsigned)nt, CSR_BLOCK, 0, stream>>>(a);
#else
    int64_t grid = (int64_t)resident_ctas((const void*)csr_pipe_kernel<T>);
    if (grid > nt) grid = nt;
    csr_pipe_kernel<T><<<(unsigned)grid, CSR_BLOCK, 0, stream>>>(a, (int)nt);
#endif
    return (int)cudaGetLastError();
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200spmv_csr_workspace_bytes(int64_t rows, int64_t nnz) {
    if (rows < 0 || nnz < 0) return 0;
    return plan_layout(csr_num_tiles(rows, nnz), nullptr, nullptr);
}

int64_t b200spmv_csr_num_tiles(int64_t rows, int64_t nnz) { return csr_num_tiles(rows, nnz); }

void b200spmv_csr_plan_params(int32_t* tile_items, int32_t* long_row, int32_t* block_threads) {
    if (tile_items) *tile_items = CSR_TILE_ITEMS;
    if (long_row) *long_row = CSR_LONG_ROW;
    if (block_threads) *block_threads = CSR_BLOCK;
}

size_t b200spmv_csr_plan_tiles_offset(void) { return PLAN_HEADER_BYTES; }

int b200spmv_csr_analyze(void* stream, int64_t rows, int64_t nnz, const void* row_offsets, int32_t base,
                         void* workspace) {
    if (rows < 0 || nnz < 0 || rows > INT32_MAX - 1 || nnz > INT32_MAX - 1 || !workspace || (rows > 0 && !row_offsets))
        return -1;
    const int64_t nt = csr_num_tiles(rows, nnz);
    PlanView v;
    plan_layout(nt, workspace, &v);
    if (rows == 0) return 0;  // (the first PLAN_HEADER_BYTES of the workspace are reserved, unused on device)
    cudaError_t e = cudaMemsetAsync(v.ctl, 0, 64, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    const int threads = 128;
    const unsigned blocks = (unsigned)((nt + 1 + threads - 1) / threads);
    csr_partition_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>((const int*)row_offsets, base, rows, nnz, nt, v);
    return (int)cudaGetLastError();
}

int b200spmv_csr_mv(void* stream, int dtype, int64_t rows, int64_t cols, int64_t nnz, const void* row_offsets,
                    const void* col_ind, const void* values, int32_t base, const void* alpha, const void* beta,
                    int scalars_on_device, const void* x, void* y, void* workspace) {
    if (rows < 0 || cols < 0 || nnz < 0 || !alpha || !beta) return -1;
    if (rows == 0) return 0;
    if (!row_offsets || !y || !workspace || (nnz > 0 && (!col_ind || !values || !x))) return -1;
    if (dtype == 0)
        return launch_csr<float>((cudaStream_t)stream, rows, nnz, row_offsets, col_ind, values, base, alpha, beta,
                                 scalars_on_device, x, y, workspace);
    if (dtype == 1)
        return launch_csr<double>((cudaStream_t)stream, rows, nnz, row_offsets, col_ind, values, base, alpha, beta,
                                  scalars_on_device, x, y, workspace);
    return -1;
}

}  // extern "C"
