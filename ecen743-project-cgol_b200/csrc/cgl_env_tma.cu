// cgl_env_tma.cu -- persistent, TMA-pipelined env step for sm_100a (side 32 / 64 / 128).
//
// Same semantics as env_step_fused_kernel (cgl_env.cu; reference: kernel `run`,
// /root/reference/CGL/CGL.py:147-181 + toggle_state :322-328 + reward :255-256), different
// data movement.  The one-CTA-per-env kernel is latency-bound: every CTA pays launch -> load ->
// barrier -> compute -> store in sequence (measured ~3 us fixed per CTA on B200).  Here
//
//   * the grid is persistent: 148 SMs x CTAS_PER_SM CTAs, each walking units u, u+G, u+2G, ...
//   * a unit is 16 KiB of stability + 2 KiB of packed world (16384 cells: 1 env of 128^2,
//     4 envs of 64^2, 16 envs of 32^2 -- contiguous in HBM);
//   * units move with the bulk-copy engine (cp.async.bulk, SASS UBLKCP): global -> shared with an
//     mbarrier transaction count, shared -> global as a bulk group.  The load of unit i+1 is in
//     flight while unit i is computed, the store of unit i-1 drains meanwhile -- no registers and
//     no LSU instructions are spent on HBM traffic;
//   * compute works in shared memory: phase B (generation, bit-sliced LOP3) reads the world tile,
//     phase C updates the stability tile in place (LDS.128 / STS.128, conflict-free).
#include "cgl_internal.cuh"

namespace cgl {

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0).
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global bulk copy, tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds_u32_tma(uint32_t shared_addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr));
    return v;
}

// ---- configuration ------------------------------------------------------------------------
constexpr int UNIT_CELLS = 16384;
constexpr int UNIT_WORDS = UNIT_CELLS / 32;      // 512 packed words  (2 KiB)
constexpr int UNIT_CHUNKS = UNIT_CELLS / 16;     // 1024 uint4 stability chunks (16 KiB)

template <int S, int THREADS>
struct TmaCfg {
    static constexpr int W = S / 32;
    static constexpr int WPE = S * W;                    // words per env
    static constexpr int SIZE = S * S;
    static constexpr int EPU = UNIT_CELLS / SIZE;        // envs per unit
    static constexpr int CPE = SIZE / 16;                // chunks per env
    static constexpr int WPT = UNIT_WORDS / THREADS;     // words per thread (phase B)
    static constexpr int CPT = UNIT_CHUNKS / THREADS;    // chunks per thread (phase C)
    // consecutive chunk slots j of one thread that fall in the same env
    static constexpr int JS = (CPE / THREADS) >= 1 ? ((CPE / THREADS) < CPT ? (CPE / THREADS) : CPT) : 1;
    static constexpr int NACC = CPT / JS;
    // shared memory map (offsets from the 4 KB aligned base)
    static constexpr int OFF_TABLES = 0;                                  // 4096
    static constexpr int OFF_STAB = 4096;                                 // 2 x 16384
    static constexpr int OFF_WORLD = OFF_STAB + 2 * UNIT_CELLS;           // 2 x 2048
    static constexpr int OFF_NEXT = OFF_WORLD + 2 * UNIT_WORDS * 4;       // 2048
    static constexpr int OFF_MIX = OFF_NEXT + UNIT_WORDS * 4;             // 4096
    static constexpr int OFF_RED = OFF_MIX + UNIT_WORDS * 8;              // 2 x 16 ints
    static constexpr int OFF_BAR = OFF_RED + 2 * 16 * 4;                  // 2 mbarriers
    static constexpr int SMEM = OFF_BAR + 16 + 4096;                      // + alignment slack
    static_assert(SIZE <= UNIT_CELLS && UNIT_CELLS % SIZE == 0 && S % 32 == 0, "unsupported side");
    static_assert(UNIT_WORDS % THREADS == 0 && UNIT_CHUNKS % THREADS == 0 && CPT % JS == 0, "bad THREADS");
};

template <int S, int THREADS>
__global__ void __launch_bounds__(THREADS)
env_step_tma_kernel(const uint32_t *__restrict__ world_in, uint32_t *__restrict__ world_out,
                    int8_t *__restrict__ stable, uint32_t n_envs, const int32_t *__restrict__ actions,
                    uint32_t spawn4, uint32_t max4, int32_t *__restrict__ reward_out,
                    uint32_t *__restrict__ alive_out, int *__restrict__ err_flag)
{
    using C = TmaCfg<S, THREADS>;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const uint32_t sbase = smem_addr(smem_dyn);
    const uint32_t tb = (sbase + 4095u) & ~4095u;          // table base: see cgl_env.cu
    unsigned char *sm = smem_dyn + (tb - sbase);
    uint32_t *tables = reinterpret_cast<uint32_t *>(sm + C::OFF_TABLES);
    uint32_t *nextw = reinterpret_cast<uint32_t *>(sm + C::OFF_NEXT);
    uint32_t *mix = reinterpret_cast<uint32_t *>(sm + C::OFF_MIX);
    int *red = reinterpret_cast<int *>(sm + C::OFF_RED);    // [0..15] reward, [16..31] alive
    const uint32_t bar0 = tb + C::OFF_BAR;

    const int t = threadIdx.x;
    const uint32_t n_units = (n_envs + C::EPU - 1) / C::EPU;

    for (int i = t; i < 1024; i += THREADS) {
        const uint32_t m = nibble_to_bytemask((uint32_t)i >> 6);
        tables[i] = (i & 32) ? (m & spawn4) : m;
    }
    if (t < 32) red[t] = 0;
    if (t == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer (thread 0): one mbarrier phase per unit load
    auto issue_load = [&](uint32_t u, int stage) {
        const uint32_t e0 = u * C::EPU;
        const uint32_t nv = (n_envs - e0) < (uint32_t)C::EPU ? (n_envs - e0) : (uint32_t)C::EPU;
        const uint32_t bar = bar0 + 8 * stage;
        mbar_expect_tx(bar, nv * (C::SIZE + C::WPE * 4));
        bulk_g2s(tb + C::OFF_STAB + stage * UNIT_CELLS, stable + (size_t)e0 * C::SIZE, nv * C::SIZE, bar);
        bulk_g2s(tb + C::OFF_WORLD + stage * UNIT_WORDS * 4, world_in + (size_t)e0 * C::WPE, nv * C::WPE * 4, bar);
    };

    uint32_t u = blockIdx.x;
    if (t == 0 && u < n_units) issue_load(u, 0);

    const uint32_t tbn = ((tb >> 8) & 0xffu) * 0x01010101u;
    const uint32_t lane_s = (t & 31) * 4, lane_b = lane_s + 128;

    for (uint32_t it = 0; u < n_units; u += gridDim.x, ++it) {
        const int stage = it & 1;
        const uint32_t e0 = u * C::EPU;
        const uint32_t nv = (n_envs - e0) < (uint32_t)C::EPU ? (n_envs - e0) : (uint32_t)C::EPU;
        uint32_t *wt = reinterpret_cast<uint32_t *>(sm + C::OFF_WORLD + stage * UNIT_WORDS * 4);
        unsigned char *st = sm + C::OFF_STAB + stage * UNIT_CELLS;

        if (t == 0) {
            // the previous unit's bulk stores must have finished READING shared memory before the
            // other stage / the nextw tile are overwritten
            bulk_wait_read0();
            const uint32_t un = u + gridDim.x;
            if (un < n_units) issue_load(un, stage ^ 1);
        }
        mbar_wait(bar0 + 8 * stage, (it >> 1) & 1);

        // ---- toggle_state(action) for every env of the unit (CGL/main.py:66, CGL.py:322-328) ----
        if (actions != nullptr && t < (int)nv) {
            const int a = actions[e0 + t];
            if (a >= 0 && a < C::SIZE) {
                wt[t * C::WPE + (a >> 5)] ^= 1u << (a & 31);          // cell index == bit index
                reinterpret_cast<int8_t *>(st)[t * C::SIZE + a] = (int8_t)spawn4;
            } else if (a != C::SIZE && err_flag != nullptr) {
                atomicOr(err_flag, 1);
            }
        }
        __syncthreads();

        // ---- phase B: next generation of every env in the unit ------------------------------
#pragma unroll
        for (int j = 0; j < C::WPT; ++j) {
            const int i = t + j * THREADS;
            const int eb = (i / C::WPE) * C::WPE, li = i % C::WPE;   // env base word, word within env
            const int r = li / C::W, w = li % C::W;
            const uint32_t *cur = wt + eb;
            const int ru = (r == 0 ? S - 1 : r - 1) * C::W, rc = r * C::W, rd = (r == S - 1 ? 0 : r + 1) * C::W;
            const int wl = (w == 0 ? C::W - 1 : w - 1), wr = (w == C::W - 1 ? 0 : w + 1);
            const uint32_t a = cur[ru + w], c = cur[rc + w], b = cur[rd + w];
            const HSum ha = hsum(west_plane(cur[ru + wl], a), a, east_plane(a, cur[ru + wr]));
            const HSum hc = hsum(west_plane(cur[rc + wl], c), c, east_plane(c, cur[rc + wr]));
            const HSum hb = hsum(west_plane(cur[rd + wl], b), b, east_plane(b, cur[rd + wr]));
            const uint32_t nxt = life_rule(ha, hc, hb, c);
            nextw[i] = nxt;
            uint32_t lo, hi;                          // byte per 4 cells: born nibble << 4 | surv nibble
            mix_nibbles(nxt & ~c, nxt & c, lo, hi);
            reinterpret_cast<uint2 *>(mix)[i] = make_uint2(lo, hi);
            if (alive_out != nullptr) {               // uniform branch
                const unsigned pop = __reduce_add_sync(0xffffffffu, (unsigned)__popc(nxt));
                if ((t & 31) == 0) atomicAdd(reinterpret_cast<unsigned *>(&red[16 + i / C::WPE]), pop);
            }
        }
        __syncthreads();

        // ---- phase C: stability tile updated in place + reward ------------------------------
        uint4 *st4 = reinterpret_cast<uint4 *>(st);
#pragma unroll
        for (int g = 0; g < C::NACC; ++g) {
            int acc = 0;
#pragma unroll
            for (int jj = 0; jj < C::JS; ++jj) {
                const int q = t + (g * C::JS + jj) * THREADS;
                const uint4 v = st4[q];
                uint32_t s[4] = {v.x, v.y, v.z, v.w};
                const uint32_t m = mix[q];
                const uint32_t sv = (m & 0x0f0f0f0fu) | tbn;
                const uint32_t bn = ((m >> 4) & 0x0f0f0f0fu) | tbn;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t surv_mask = lds_u32_tma(__byte_perm(sv, lane_s, 0x5504 + 16 * k));
                    const uint32_t born_spawn = lds_u32_tma(__byte_perm(bn, lane_b, 0x5504 + 16 * k));
                    s[k] = stable_update4(s[k], surv_mask, born_spawn, max4);
                    acc = __dp4a((int)s[k], 0x01010101, acc);
                }
                st4[q] = make_uint4(s[0], s[1], s[2], s[3]);
            }
            if (reward_out != nullptr) {              // uniform branch
                acc = __reduce_add_sync(0xffffffffu, acc);
                if ((t & 31) == 0) atomicAdd(&red[(t + g * C::JS * THREADS) / C::CPE], acc);
            }
        }
        fence_async_smem();                           // generic-proxy writes -> visible to the bulk engine
        __syncthreads();

        if (t == 0) {
            bulk_s2g(stable + (size_t)e0 * C::SIZE, tb + C::OFF_STAB + stage * UNIT_CELLS, nv * C::SIZE);
            bulk_s2g(world_out + (size_t)e0 * C::WPE, tb + C::OFF_NEXT, nv * C::WPE * 4);
            bulk_commit();
        }
        if (t < 16) {                                 // the reader of a slot is also the one that clears it
            const int rw = red[t], al = red[16 + t];
            red[t] = 0;
            red[16 + t] = 0;
            if (t < (int)nv) {
                if (reward_out != nullptr) reward_out[e0 + t] = rw;
                if (alive_out != nullptr) alive_out[e0 + t] = (uint32_t)al;
            }
        }
    }
    if (t == 0) bulk_wait_all0();                     // shared memory must outlive the bulk stores
}

template <int S, int THREADS>
static int launch_env_tma(const uint32_t *win, uint32_t *wout, int8_t *stable, uint64_t n_envs,
                          const int32_t *actions, int spawn, int stable_max, int32_t *reward,
                          uint32_t *alive, int *err, cudaStream_t st)
{
    using C = TmaCfg<S, THREADS>;
    static int ctas_per_sm = 0;
    static PerDeviceOnce once;
    if (once.first()) {
        CGL_CUDA(cudaFuncSetAttribute(env_step_tma_kernel<S, THREADS>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int n = 0;
        CGL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, env_step_tma_kernel<S, THREADS>, THREADS, C::SMEM));
        ctas_per_sm = n > 0 ? n : 1;
    }
    const uint64_t n_units = (n_envs + C::EPU - 1) / C::EPU;
    uint64_t grid = (uint64_t)sm_count() * ctas_per_sm;      // persistent: one resident wave
    if (grid > n_units) grid = n_units;
    env_step_tma_kernel<S, THREADS><<<(unsigned)grid, THREADS, C::SMEM, st>>>(
        win, wout, stable, (uint32_t)n_envs, actions, rep4(spawn), rep4(stable_max), reward, alive, err);
    CGL_LAUNCH_CHECK();
    return 0;
}

}  // namespace cgl

using namespace cgl;

// Returns -100 if `side` has no TMA instantiation (caller falls back to the per-env kernel).
extern "C" int cgl_env_step_tma(uint32_t *win, uint32_t *wout, int8_t *stable, uint64_t n_envs, uint32_t side,
                                const int32_t *actions, int spawn, int stable_max, int32_t *reward,
                                uint32_t *alive, int *err, cgl_stream_t stream, int threads)
{
    cudaStream_t st = as_stream(stream);
#define CGL_TMA_CASE(S)                                                                                        \
    case S:                                                                                                    \
        return threads == 128 ? launch_env_tma<S, 128>(win, wout, stable, n_envs, actions, spawn, stable_max,  \
                                                       reward, alive, err, st)                                 \
                              : launch_env_tma<S, 256>(win, wout, stable, n_envs, actions, spawn, stable_max,  \
                                                       reward, alive, err, st)
    switch (side) {
        CGL_TMA_CASE(32);
        CGL_TMA_CASE(64);
        CGL_TMA_CASE(128);
    }
#undef CGL_TMA_CASE
    return -100;
}
