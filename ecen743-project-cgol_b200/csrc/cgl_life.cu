// cgl_life.cu -- world-only generations ("life mode") of large bit-packed grids for sm_100a.
//
// Reference semantics: the world half of kernel `run`, /root/reference/CGL/CGL.py:154-170
// (torus neighbour count + B3/S23).  The reference cannot even express these sizes (its
// `unsigned int` index math overflows above side 46340).
//
// life_rows_kernel (k = 1 streaming):  HBM traffic = 1 bit read + 1 bit written per cell
// (0.25 B / cell-update).  A thread owns a 128-cell-wide column (one uint4 per row) and walks
// down RPT rows keeping a 3-row window of horizontal partial sums in registers; horizontal
// neighbours come from the adjacent lanes by warp shuffle, the two warp-edge lanes load one halo
// word.  Per 32 cells: 2 SHF + 4 LOP3 (row sums, reused by 3 output rows) + 6 LOP3 (rule).
#include "cgl_internal.cuh"

namespace cgl {

struct RowRaw { uint4 v; uint32_t halo; };
struct RowSums { HSum h[4]; uint4 c; };

// Load one row segment of this lane (all zero for a dead row outside an open band).
__device__ __forceinline__ RowRaw load_row_raw(const uint32_t *__restrict__ grid, int64_t r,
                                               uint32_t rows, uint32_t W, int wrap_rows,
                                               uint32_t c4, uint32_t hidx, bool valid, bool edge)
{
    RowRaw o;
    o.v = make_uint4(0, 0, 0, 0);
    o.halo = 0;
    if (r < 0 || r >= (int64_t)rows) {
        if (!wrap_rows) return o;
        r = r < 0 ? r + rows : r - rows;
    }
    const uint32_t *row = grid + (uint64_t)r * W;
    if (valid) o.v = __ldg(reinterpret_cast<const uint4 *>(row) + c4);
    if (edge) o.halo = __ldg(row + hidx);
    return o;
}

__device__ __forceinline__ RowSums row_sums(const RowRaw &raw, bool edge_l, bool edge_r)
{
    uint32_t left = __shfl_up_sync(0xffffffffu, raw.v.w, 1);
    uint32_t right = __shfl_down_sync(0xffffffffu, raw.v.x, 1);
    if (edge_l) left = raw.halo;
    if (edge_r) right = raw.halo;
    RowSums s;
    s.c = raw.v;
    s.h[0] = hsum(west_plane(left, raw.v.x), raw.v.x, east_plane(raw.v.x, raw.v.y));
    s.h[1] = hsum(west_plane(raw.v.x, raw.v.y), raw.v.y, east_plane(raw.v.y, raw.v.z));
    s.h[2] = hsum(west_plane(raw.v.y, raw.v.z), raw.v.z, east_plane(raw.v.z, raw.v.w));
    s.h[3] = hsum(west_plane(raw.v.z, raw.v.w), raw.v.w, east_plane(raw.v.w, right));
    return s;
}

constexpr int LIFE_ROWS_THREADS = 128;   // 4 warps = 4 horizontally adjacent column groups
constexpr int LIFE_ROWS_PF = 4;          // rows loaded ahead per thread

template <bool COUNT>
__global__ void __launch_bounds__(LIFE_ROWS_THREADS)
life_rows_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n_envs,
                 uint32_t rows, uint32_t W, uint32_t rpt, int wrap_rows, uint32_t n_cgroups,
                 uint32_t n_rblocks, uint32_t *__restrict__ alive_out)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (LIFE_ROWS_THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t cg = (uint32_t)(warp % n_cgroups);
    const uint64_t tmp = warp / n_cgroups;
    const uint32_t rb = (uint32_t)(tmp % n_rblocks);
    const uint64_t e = tmp / n_rblocks;
    if (e >= n_envs) return;                                   // warp-uniform

    const uint32_t W4 = W >> 2;
    const uint32_t c4 = cg * 32 + lane;
    const bool valid = c4 < W4;
    const bool edge_l = lane == 0;
    const bool edge_r = valid && (lane == 31 || c4 == W4 - 1);
    const uint32_t hidx = edge_l ? (c4 == 0 ? W - 1 : 4 * c4 - 1) : (c4 == W4 - 1 ? 0 : 4 * c4 + 4);
    const bool edge = valid && (edge_l || edge_r);
    const uint32_t *gin = in + e * (uint64_t)rows * W;
    uint32_t *gout = out + e * (uint64_t)rows * W;

    const int64_t r0 = (int64_t)rb * rpt;
    const int64_t r1 = (r0 + rpt < (int64_t)rows) ? r0 + rpt : (int64_t)rows;

    RowSums up = row_sums(load_row_raw(gin, r0 - 1, rows, W, wrap_rows, c4, hidx, valid, edge), edge_l, edge_r);
    RowSums mid = row_sums(load_row_raw(gin, r0, rows, W, wrap_rows, c4, hidx, valid, edge), edge_l, edge_r);
    unsigned pop = 0;

    for (int64_t r = r0; r < r1; r += LIFE_ROWS_PF) {
        RowRaw raw[LIFE_ROWS_PF];
#pragma unroll
        for (int u = 0; u < LIFE_ROWS_PF; ++u) {
            raw[u].v = make_uint4(0, 0, 0, 0);
            raw[u].halo = 0;
            if (r + u < r1) raw[u] = load_row_raw(gin, r + u + 1, rows, W, wrap_rows, c4, hidx, valid, edge);
        }
#pragma unroll
        for (int u = 0; u < LIFE_ROWS_PF; ++u) {
            if (r + u < r1) {                                   // warp-uniform
                const RowSums dn = row_sums(raw[u], edge_l, edge_r);
                uint4 o;
                o.x = life_rule(up.h[0], mid.h[0], dn.h[0], mid.c.x);
                o.y = life_rule(up.h[1], mid.h[1], dn.h[1], mid.c.y);
                o.z = life_rule(up.h[2], mid.h[2], dn.h[2], mid.c.z);
                o.w = life_rule(up.h[3], mid.h[3], dn.h[3], mid.c.w);
                if (valid) reinterpret_cast<uint4 *>(gout + (uint64_t)(r + u) * W)[c4] = o;
                if (COUNT) pop += __popc(o.x) + __popc(o.y) + __popc(o.z) + __popc(o.w);
                up = mid;
                mid = dn;
            }
        }
    }
    if (COUNT) {
        if (!valid) pop = 0;
        pop = __reduce_add_sync(0xffffffffu, pop);
        if (lane == 0 && pop) atomicAdd(alive_out + e, pop);
    }
}

static uint32_t pick_rows_per_strip(uint64_t n_envs, uint32_t rows, uint32_t n_cgroups)
{
    // aim for >= 8 strips per resident warp slot (148 SMs x 16 warps) to keep the tail small,
    // but never fewer than 16 rows per strip (2 halo rows are re-read per strip).
    const uint64_t want = (uint64_t)sm_count() * 16 * 8;
    uint64_t rpt = 128;
    while (rpt > 16 && n_envs * n_cgroups * ((rows + rpt - 1) / rpt) < want) rpt >>= 1;
    return (uint32_t)rpt;
}

}  // namespace cgl

using namespace cgl;

extern "C" int cgl_life_step_generic(const uint32_t *in, uint32_t *out, uint64_t n_envs, uint32_t rows,
                                     uint32_t cols, int wrap_rows, uint32_t *alive_out,
                                     cgl_stream_t stream);

extern "C" int cgl_life_step(const uint32_t *in, uint32_t *out, uint64_t n_envs, uint32_t rows,
                             uint32_t cols, int wrap_rows, uint32_t *alive_out, cgl_stream_t stream)
{
    CGL_REQUIRE(in && out && n_envs && rows && cols && in != out, CGL_E_BADARG,
                "cgl_life_step: bad argument");
    if (cols % 128 != 0 || cols < 2048 || (cols / 128) % 32 == 1 || n_envs >= (1ull << 31))
        return cgl_life_step_generic(in, out, n_envs, rows, cols, wrap_rows, alive_out, stream);
    cudaStream_t st = as_stream(stream);
    const uint32_t W = cols / 32, W4 = W / 4;
    const uint32_t n_cgroups = (W4 + 31) / 32;
    const uint32_t rpt = pick_rows_per_strip(n_envs, rows, n_cgroups);
    const uint32_t n_rblocks = (rows + rpt - 1) / rpt;
    const uint64_t warps = n_envs * n_cgroups * n_rblocks;
    const uint64_t blocks = (warps + (LIFE_ROWS_THREADS / 32) - 1) / (LIFE_ROWS_THREADS / 32);
    CGL_REQUIRE(blocks < (1ull << 31), CGL_E_BADARG, "cgl_life_step: grid too large");
    if (alive_out != nullptr) {
        CGL_CUDA(cudaMemsetAsync(alive_out, 0, n_envs * sizeof(uint32_t), st));
        life_rows_kernel<true><<<(unsigned)blocks, LIFE_ROWS_THREADS, 0, st>>>(
            in, out, (uint32_t)n_envs, rows, W, rpt, wrap_rows, n_cgroups, n_rblocks, alive_out);
    } else {
        life_rows_kernel<false><<<(unsigned)blocks, LIFE_ROWS_THREADS, 0, st>>>(
            in, out, (uint32_t)n_envs, rows, W, rpt, wrap_rows, n_cgroups, n_rblocks, nullptr);
    }
    CGL_LAUNCH_CHECK();
    return 0;
}

extern "C" int cgl_life_run(uint32_t *buf_a, uint32_t *buf_b, uint32_t rows, uint32_t cols,
                            int wrap_rows, uint32_t gens, uint32_t k, int *result_in_a_out,
                            cgl_stream_t stream)
{
    CGL_REQUIRE(buf_a && buf_b && rows && cols && buf_a != buf_b, CGL_E_BADARG, "cgl_life_run: bad argument");
    (void)k;   // temporal blocking lands in cgl_life_tb.cu; k = 1 streams
    uint32_t *src = buf_a, *dst = buf_b;
    for (uint32_t g = 0; g < gens; ++g) {
        int rc = cgl_life_step(src, dst, 1, rows, cols, wrap_rows, nullptr, stream);
        if (rc) return rc;
        uint32_t *t = src; src = dst; dst = t;
    }
    if (result_in_a_out) *result_in_a_out = (src == buf_a) ? 1 : 0;
    return 0;
}
