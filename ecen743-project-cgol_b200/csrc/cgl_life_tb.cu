// cgl_life_tb.cu -- temporal blocking for life mode: K generations per launch, one HBM pass.
//
// Reference semantics: K applications of the world half of kernel `run`
// (/root/reference/CGL/CGL.py:154-170).  BASELINE.json configs[4].
//
// Time-skewed register pipeline (no shared memory, no re-reads):
//   * a warp owns a strip of 30 packed word-columns (+1 halo word-column on each side, lanes 0 and
//     31) and streams down the rows; lane l holds word-column 30*cg + l - 1.  The halo words go
//     stale by one bit-column per generation, so lanes 1..30 stay exact for K <= 32.
//   * level g (1..K) keeps a 2-row window of horizontal partial sums of generation g-1; when row
//     rho of generation g-1 arrives it emits row rho-1 of generation g, which feeds level g+1 on
//     the NEXT row step (one step of skew makes the K levels independent within a step: ILP = K).
//   * per 32 cells and generation: 2 SHFL + 2 SHF + 10 LOP3 -- the k = 1 kernel's ALU work -- while
//     HBM traffic drops to (32/30 read + 1 write) bits per cell per K generations.
// Row r of generation K leaves the pipeline 3K-1 row steps after gen-0 row r-K entered it, so a
// strip of L rows costs L + 3K - 1 steps.
#include <stdlib.h>

#include "cgl_internal.cuh"

namespace cgl {

struct Win {            // window of one level: rows (a-1, a) of the level's input generation
    uint32_t us0, us1;                  // row a-1: west+centre+east sum bits
    uint32_t ms0, ms1, mt0, mt1, mc;    // row a: sums with / without the centre, and the centre word
};

constexpr int TB_THREADS = 128;
constexpr int TB_UNROLL = 6;            // row steps per loop trip (loads issued up front)
constexpr int TB_COLS = 30;             // valid word-columns per warp

template <int K>
__global__ void __launch_bounds__(TB_THREADS)
life_tb_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t rows, uint32_t W,
               uint32_t rpt, int wrap_rows, uint32_t n_cgroups, uint32_t n_rblocks)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (TB_THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t cg = warp % n_cgroups;
    uint32_t rb = warp / n_cgroups;
    // no early exit: padding warps redo the last strip with stores off, so that every shuffle
    // below is provably convergent (plain SHFL, no divergence fallback path)
    const bool warp_ok = rb < n_rblocks;
    rb = warp_ok ? rb : n_rblocks - 1;

    const int wi = (int)(cg * TB_COLS + lane) - 1;             // word column of this lane (may be -1 or >= W)
    const uint32_t wcol = wi < 0 ? (uint32_t)(wi + (int)W) : ((uint32_t)wi >= W ? (uint32_t)wi - W : (uint32_t)wi);
    const bool store_ok = warp_ok && lane >= 1 && lane <= TB_COLS && (uint32_t)wi < W;

    const int r0 = (int)(rb * rpt);
    const int r1 = (r0 + (int)rpt < (int)rows) ? r0 + (int)rpt : (int)rows;
    const int irows = (int)rows;
    const int rstart = r0 - K;                                  // first gen-0 row fed to level 1 (> -rows)
    const int n_steps = (int)rpt + 3 * K - 1;                   // same trip count for every warp (uniform loop)
    // Step s loads gen-0 row rstart + s.  It is needed (and inside the grid, for open rows) iff
    // s_lo <= s < s_hi; rows from r1 + K on can no longer reach an output row of this strip.
    const int s_lo = wrap_rows ? 0 : (rstart < 0 ? -rstart : 0);
    const int last = (wrap_rows || r1 + K < irows) ? r1 + K : irows;
    const uint32_t span = (uint32_t)(last - rstart - s_lo);
    const uint32_t out_rows = (uint32_t)(r1 - r0);
    int rw = rstart < 0 ? rstart + irows : rstart;              // row index modulo rows (only used where loads are on)
    const uint32_t *ip = in + wcol;
    uint32_t *op = out + (store_ok ? (uint32_t)wi : 0u);

    Win win[K];
    uint32_t pend[K];                   // pend[g] = output of level g+1 at the previous step
#pragma unroll
    for (int g = 0; g < K; ++g) {
        win[g] = Win{0, 0, 0, 0, 0, 0, 0};
        pend[g] = 0;
    }

    for (int s0 = 0; s0 < n_steps; s0 += TB_UNROLL) {
        uint32_t raw[TB_UNROLL];
#pragma unroll
        for (int u = 0; u < TB_UNROLL; ++u) {
            const uint32_t ru = umin((uint32_t)rw + u, (uint32_t)rw + u - rows);      // (rw + u) mod rows
            raw[u] = 0;
            if ((uint32_t)(s0 + u - s_lo) < span) raw[u] = __ldg(ip + ru * W);          // rows * W < 2^32
        }
        rw += TB_UNROLL;
        rw = rw >= irows ? rw - irows : rw;
#pragma unroll
        for (int u = 0; u < TB_UNROLL; ++u) {
            // levels in descending order: level g+1 reads pend[g-1] before level g overwrites it
#pragma unroll
            for (int g = K - 1; g >= 0; --g) {
                const uint32_t x = (g == 0) ? raw[u] : pend[g - 1];
                const uint32_t left = __shfl_up_sync(0xffffffffu, x, 1);
                const uint32_t right = __shfl_down_sync(0xffffffffu, x, 1);
                const HSum d = hsum(west_plane(left, x), x, east_plane(x, right));
                const HSum up = {win[g].us0, win[g].us1, 0, 0};
                const HSum mid = {0, 0, win[g].mt0, win[g].mt1};
                pend[g] = life_rule(up, mid, d, win[g].mc);
                win[g].us0 = win[g].ms0; win[g].us1 = win[g].ms1;
                win[g].ms0 = d.s0; win[g].ms1 = d.s1; win[g].mt0 = d.t0; win[g].mt1 = d.t1; win[g].mc = x;
            }
            // row r0 + so of generation K has just left level K
            const uint32_t so = (uint32_t)(s0 + u - (3 * K - 1));
            if (store_ok && so < out_rows) op[((uint32_t)r0 + so) * W] = pend[K - 1];
        }
    }
}

template <int K>
static int launch_tb(const uint32_t *in, uint32_t *out, uint32_t rows, uint32_t cols, int wrap_rows,
                     cudaStream_t st)
{
    static int rpt_knob = -1;
    if (rpt_knob < 0) {
        const char *e = getenv("CGL_TB_ROWS");
        rpt_knob = e ? atoi(e) : 0;
    }
    const uint32_t W = cols / 32;
    const uint32_t n_cgroups = (W + TB_COLS - 1) / TB_COLS;
    // strip length: long enough to amortise the 3K-1 fill steps, short enough for >= ~4 waves of warps
    uint32_t rpt = rpt_knob > 0 ? (uint32_t)rpt_knob : 64u * K;
    if (rpt < 64) rpt = 64;
    const uint64_t want = (uint64_t)sm_count() * 16 * 4;
    while (rpt > 24u * K && rpt > 32 && (uint64_t)n_cgroups * ((rows + rpt - 1) / rpt) < want && rpt_knob <= 0) rpt >>= 1;
    if (rpt > rows) rpt = rows;
    const uint32_t n_rblocks = (rows + rpt - 1) / rpt;
    const uint64_t warps = (uint64_t)n_cgroups * n_rblocks;
    const uint64_t blocks = (warps + (TB_THREADS / 32) - 1) / (TB_THREADS / 32);
    CGL_REQUIRE(blocks < (1ull << 31) && rows < (1u << 30) && (uint64_t)rows * W < (1ull << 32), CGL_E_BADARG,
                "cgl_life_run: grid too large for the k-blocked kernel (rows * cols/32 must be < 2^32)");
    life_tb_kernel<K><<<(unsigned)blocks, TB_THREADS, 0, st>>>(in, out, rows, W, rpt, wrap_rows, n_cgroups, n_rblocks);
    CGL_LAUNCH_CHECK();
    return 0;
}

static int tb_dispatch(int k, const uint32_t *in, uint32_t *out, uint32_t rows, uint32_t cols, int wrap_rows,
                       cudaStream_t st)
{
    switch (k) {
        case 1: return launch_tb<1>(in, out, rows, cols, wrap_rows, st);
        case 2: return launch_tb<2>(in, out, rows, cols, wrap_rows, st);
        case 3: return launch_tb<3>(in, out, rows, cols, wrap_rows, st);
        case 4: return launch_tb<4>(in, out, rows, cols, wrap_rows, st);
        case 6: return launch_tb<6>(in, out, rows, cols, wrap_rows, st);
        case 8: return launch_tb<8>(in, out, rows, cols, wrap_rows, st);
        case 12: return launch_tb<12>(in, out, rows, cols, wrap_rows, st);
        case 16: return launch_tb<16>(in, out, rows, cols, wrap_rows, st);
    }
    set_error("cgl_life_run: no temporal-blocking kernel for k=%d", k);
    return CGL_E_BADARG;
}

}  // namespace cgl

using namespace cgl;

extern "C" int cgl_life_step(const uint32_t *in, uint32_t *out, uint64_t n_envs, uint32_t rows,
                             uint32_t cols, int wrap_rows, uint32_t *alive_out, cgl_stream_t stream);

// k generations per launch where a kernel exists for the block size, else smaller blocks.
// With wrap_rows = 0 the k rows next to each open edge are a ghost zone: their contents after a
// k-blocked launch are unspecified (they differ from k single steps); every row at least k rows
// away from an open edge is exact.
extern "C" int cgl_life_run(uint32_t *buf_a, uint32_t *buf_b, uint32_t rows, uint32_t cols,
                            int wrap_rows, uint32_t gens, uint32_t k, int *result_in_a_out,
                            cgl_stream_t stream)
{
    CGL_REQUIRE(buf_a && buf_b && rows && cols && buf_a != buf_b, CGL_E_BADARG, "cgl_life_run: bad argument");
    CGL_REQUIRE(k >= 1 && k <= 16, CGL_E_BADARG, "cgl_life_run: k must be in 1..16");
    cudaStream_t st = as_stream(stream);
    static const int sizes[] = {16, 12, 8, 6, 4, 3, 2, 1};
    uint32_t *src = buf_a, *dst = buf_b;
    uint32_t left = gens;
    const bool tiled = cols % 32 == 0 && cols >= 32 * TB_COLS && rows >= 8;
    while (left > 0) {
        int step = 1;
        if (tiled && k > 1) {
            for (int s : sizes)
                if ((uint32_t)s <= k && (uint32_t)s <= left) { step = s; break; }
        }
        int rc;
        if (step == 1) rc = cgl_life_step(src, dst, 1, rows, cols, wrap_rows, nullptr, stream);
        else rc = tb_dispatch(step, src, dst, rows, cols, wrap_rows, st);
        if (rc) return rc;
        uint32_t *t = src; src = dst; dst = t;
        left -= (uint32_t)step;
    }
    if (result_in_a_out) *result_in_a_out = (src == buf_a) ? 1 : 0;
    return 0;
}
