// cgl_sim1.cu -- the single-environment step behind the `CGL.sim` facade: ONE launch per
// toggle_state -> step -> get_stable -> reward round of the reference's DQN loop.
//
// Reference: the body of the training loop, /root/reference/CGL/main.py:64-72 --
//   env.toggle_state(action)   CGL/CGL.py:322-328   (fork: CGL_action+/CGL.py:377-384, masked write)
//   env.step()                 CGL/CGL.py:247-252, kernel `run` :147-181 + four PCIe copies :203-208
//   env.get_stable(shallow)    CGL/CGL.py:281-285   (the live int8 buffer)
//   env.reward()               CGL/CGL.py:255-256
// On one environment a step is latency, not bandwidth: what counts is how many launches, copies and
// synchronisations sit between the caller's action and the observation it gets back.  Here:
//   * the action arrives BY VALUE (a kernel argument): nothing the host rewrites later is read by the kernel,
//     so back-to-back toggle+step pairs need no synchronisation between them;
//   * the new stability plane is stored twice -- to the device plane and straight into the caller's pinned,
//     host-mapped observation mirror (posted PCIe writes) -- so there is no device-to-host copy;
//   * reward and live count are reduced in the same kernel and written to a host-mapped result block, followed
//     (system-scope fences in between) by the caller's sequence number: the host polls that word instead of
//     synchronising the stream.
// One CTA owns the environment: world bits unpacked to one byte per cell in shared memory, any side up to
// CGL_SIM1_MAX_SIDE, every dead-cell rule of cgl_bits.cuh.
#include <stdlib.h>

#include "cgl_internal.cuh"

namespace cgl {

constexpr uint32_t SIM1_MAX_SIDE = 448;          // one byte plane + the packed next world: 448^2 + 448 * 56 B of shared memory

// Thread layout: a thread serves FOUR consecutive cells of one row (the last thread of a row fewer); TPR threads per
// row, blockDim / TPR rows per pass.  One division per thread, none per cell.
__global__ void __launch_bounds__(1024)
sim1_step_kernel(const uint32_t *__restrict__ world_in, uint32_t *__restrict__ world_out, int8_t *stable,
                 uint32_t side, uint32_t W, uint32_t action, int8_t spawn, int8_t stable_max, int rule, int8_t empty,
                 int8_t empty_min, int masked, int8_t *obs_mirror, int32_t *result, uint32_t seq, uint32_t tpr,
                 uint32_t rows_per_pass)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const uint32_t size = side * side, n_words = side * W;
    uint8_t *a = smem_dyn;                                        // current world, one byte per cell
    uint32_t *nw = reinterpret_cast<uint32_t *>(smem_dyn + ((size + 15u) & ~15u));     // next world, packed
    __shared__ int red[2];
    const uint32_t ty = threadIdx.x / tpr, x0 = (threadIdx.x - ty * tpr) * 4;
    const bool lane_ok = ty < rows_per_pass;
    const uint32_t nx = side - x0 < 4 ? side - x0 : 4;            // cells of this thread (x0 < side by construction)
    if (threadIdx.x == 0) { red[0] = 0; red[1] = 0; }
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) nw[i] = 0;
    // world bits -> bytes, the toggle applied on the way (action == size: nothing to toggle)
    if (lane_ok)
        for (uint32_t y = ty; y < side; y += rows_per_pass) {
            const uint32_t bits = world_in[y * W + (x0 >> 5)] >> (x0 & 31);        // x0 % 4 == 0: one word
            for (uint32_t k = 0; k < nx; ++k) {
                const uint32_t i = y * side + x0 + k;
                a[i] = (uint8_t)(((bits >> k) & 1u) ^ (i == action ? 1u : 0u));
            }
        }
    __syncthreads();

    int acc = 0;
    const bool vec = (side & 3u) == 0;                            // rows start 4-byte aligned: one 32-bit access
    if (lane_ok)
        for (uint32_t y = ty; y < side; y += rows_per_pass) {
            const uint32_t yu = (y == 0 ? side : y) - 1, yd = (y + 1 == side) ? 0 : y + 1;
            const uint8_t *ru = a + yu * side, *rc = a + y * side, *rd = a + yd * side;
            const uint32_t base = y * side + x0;
            uint32_t sv = 0;
            if (vec) sv = *reinterpret_cast<const uint32_t *>(stable + base);
            else for (uint32_t k = 0; k < nx; ++k) sv |= (uint32_t)(uint8_t)stable[base + k] << (8 * k);
            // column sums of the three rows for x0-1 .. x0+nx (torus columns)
            uint32_t col3[6], mid[6];
#pragma unroll
            for (uint32_t k = 0; k < 6; ++k) {
                if (k >= nx + 2) break;
                uint32_t x = x0 + k;                               // column x - 1
                x = x == 0 ? side - 1 : (x - 1 >= side ? x - 1 - side : x - 1);
                mid[k] = rc[x];
                col3[k] = ru[x] + mid[k] + rd[x];
            }
            uint32_t out = 0, nbits = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                if (k >= nx) break;
                const uint32_t p = mid[k + 1];
                const uint32_t cnt = col3[k] + col3[k + 1] + col3[k + 2] - p;
                const uint32_t q = (cnt == 3u) || (cnt == 2u && p);
                nbits |= q << k;
                int8_t s = (int8_t)(sv >> (8 * k));
                // toggle_state: SPAWN (base env: even when toggled to dead, N2; fork: 0 then)
                if (base + k == action) s = (masked && !p) ? (int8_t)0 : spawn;
                s = stable_update1_rule(rule, s, p != 0, q != 0, spawn, stable_max, empty, empty_min);
                acc += s;
                out |= (uint32_t)(uint8_t)s << (8 * k);
            }
            if (nbits) atomicOr(&nw[y * W + (x0 >> 5)], nbits << (x0 & 31));
            if (vec) {
                *reinterpret_cast<uint32_t *>(stable + base) = out;
                if (obs_mirror != nullptr) *reinterpret_cast<uint32_t *>(obs_mirror + base) = out;
            } else {
                for (uint32_t k = 0; k < nx; ++k) {
                    stable[base + k] = (int8_t)(out >> (8 * k));
                    if (obs_mirror != nullptr) obs_mirror[base + k] = (int8_t)(out >> (8 * k));
                }
            }
        }
    if (obs_mirror != nullptr) __threadfence_system();       // my mirror stores are visible to the host ...
    __syncthreads();                                          // ... before anyone publishes the sequence number

    uint32_t pop = 0;
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) {
        const uint32_t word = nw[i];
        pop += __popc(word);
        world_out[i] = word;
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    pop = __reduce_add_sync(0xffffffffu, pop);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&red[0], acc);
        atomicAdd(reinterpret_cast<unsigned *>(&red[1]), pop);
    }
    __syncthreads();
    if (threadIdx.x == 0 && result != nullptr) {
        volatile int32_t *r = result;
        r[0] = red[0];
        r[1] = red[1];
        __threadfence_system();
        r[2] = (int32_t)seq;
    }
}


// ---- the resident form: the environment stays on chip BETWEEN steps -----------------------------------------
// A launch per step costs the facade ~9 us of launch latency for ~1 us of work.  sim1_serve_kernel is launched once,
// keeps the world (one byte per cell, two planes) and the stability plane in shared memory, and then serves steps:
// thread 0 polls a 64-bit command word in pinned host memory ((seq << 32) | action, written by the host with one
// store), the CTA performs toggle + generation + stability + reward, stores the new observation into the caller's
// pinned mirror and publishes {reward, alive, seq} with ONE 16-byte store the host polls.  No launch, no copy, no
// stream synchronisation per step.  The kernel LEAVES -- device planes written back, result[4] = launch_id -- when
// the host sends SIM1_QUIT or when no command arrived for `linger_ns`: a host that goes away to do other work (a
// Q-network update, a device-wide synchronise) is never blocked for longer than that, and its next step simply
// launches the kernel again.
constexpr uint32_t SIM1_SERVE_MAX_SIDE = 256;    // two halo'd byte planes + the stability plane: 197 KB of shared memory
constexpr uint32_t SIM1_QUIT = 0xFFFFFFFEu;
constexpr uint32_t SERVE_COPY_THREADS = 32;      // threads that write (and fence) the observation mirror; CGL_SERVE_COPIERS overrides

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Shared-memory layout of a world plane: one byte per cell WITH torus halos, so that a thread's four cells and
// their neighbours are three aligned 32-bit words plus six bytes and the neighbour count is packed-byte arithmetic
// (no wrap logic, any side).  Row r (-1 .. side) starts at (r + 1) * stride; cell x sits at byte 4 + x, the left halo
// (= cell side-1) at byte 3, the right halo (= cell 0) at byte 4 + side; every other byte of a row stays 0.
__device__ __forceinline__ uint32_t serve_stride(uint32_t side) { return (side + 5u + 3u) & ~3u; }

// Write one cell and every halo image of it.
__device__ __forceinline__ void serve_set_cell(uint8_t *pl, uint32_t side, uint32_t S, uint32_t y, uint32_t x, uint8_t v)
{
    for (int ri = 0; ri < 3; ++ri) {
        if (ri == 1 && y != 0) continue;                  // image in the bottom halo row
        if (ri == 2 && y != side - 1) continue;           // image in the top halo row
        uint8_t *row = pl + (ri == 0 ? y + 1 : (ri == 1 ? side + 1 : 0)) * S;
        row[4 + x] = v;
        if (x == 0) row[4 + side] = v;
        if (x == side - 1) row[3] = v;
    }
}

__global__ void __launch_bounds__(1024)
sim1_serve_kernel(uint32_t *world, int8_t *stable, uint32_t side, uint32_t W, int8_t spawn, int8_t stable_max,
                  int rule, int8_t empty, int8_t empty_min, int masked, int8_t *obs_mirror, int32_t *result,
                  const unsigned long long *cmd, uint32_t last_seq, uint32_t launch_id, unsigned long long linger_ns,
                  uint32_t tpr, uint32_t rows_per_pass, uint32_t copy_threads)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const uint32_t size = side * side, n_words = side * W, S = serve_stride(side);
    const uint32_t plane_bytes = ((side + 2) * S + 15u) & ~15u;
    uint8_t *cur = smem_dyn, *nxt = smem_dyn + plane_bytes;
    int8_t *stab = reinterpret_cast<int8_t *>(smem_dyn + 2 * plane_bytes);
    __shared__ int red[2];
    __shared__ uint32_t bc[2];
    const uint32_t ty = threadIdx.x / tpr, x0 = (threadIdx.x - ty * tpr) * 4;
    const bool lane_ok = ty < rows_per_pass;
    const uint32_t nx = side - x0 < 4 ? side - x0 : 4;
    const bool vec = (side & 3u) == 0;
    const uint32_t valid = nx == 4 ? 0xffffffffu : ((1u << (8 * nx)) - 1u);
    const uint32_t spawn4 = rep4(spawn), max4 = rep4(stable_max), min4 = rep4(empty_min), empty4 = rep4(empty);

    for (uint32_t i = threadIdx.x; i < 2 * plane_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_dyn)[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (side + 2) * (side + 2); i += blockDim.x) {
        const uint32_t r = i / (side + 2), c = i - r * (side + 2);           // halo coordinates: cell (r - 1, c - 1)
        const uint32_t y = r == 0 ? side - 1 : (r == side + 1 ? 0 : r - 1);
        const uint32_t x = c == 0 ? side - 1 : (c == side + 1 ? 0 : c - 1);
        cur[r * S + 3 + c] = (uint8_t)((world[y * W + (x >> 5)] >> (x & 31)) & 1u);
    }
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) stab[i] = stable[i];
    uint32_t done = last_seq;
    unsigned long long t_seen = 0;                                 // (thread 0) when the current command was seen
    __syncthreads();

    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned long long c;
            for (;;) {
                c = ld_sys_u64(cmd);
                if ((uint32_t)(c >> 32) != done) break;
                if (globaltimer_ns() - t0 > linger_ns) { c = SIM1_QUIT; break; }
            }
            t_seen = globaltimer_ns();
            const uint32_t action = (uint32_t)c;
            if (action < size) {                                   // toggle_state before the step
                const uint32_t y = action / side, x = action - y * side;
                const uint8_t v = cur[(y + 1) * S + 4 + x] ^ 1u;
                serve_set_cell(cur, side, S, y, x, v);
                stab[action] = (masked && !v) ? (int8_t)0 : spawn;
            }
            bc[0] = action;
            bc[1] = (uint32_t)(c >> 32);
            red[0] = 0;
            red[1] = 0;
        }
        __syncthreads();
        if (bc[0] == SIM1_QUIT) break;
        done = bc[1];

        int acc = 0;
        uint32_t pop = 0;
        if (lane_ok)
            for (uint32_t y = ty; y < side; y += rows_per_pass) {
                const uint32_t off = (y + 1) * S + 4 + x0;         // 4-byte aligned
                const uint8_t *rm = cur + off;
                const uint32_t u = *reinterpret_cast<const uint32_t *>(rm - S);
                const uint32_t m = *reinterpret_cast<const uint32_t *>(rm);
                const uint32_t d = *reinterpret_cast<const uint32_t *>(rm + S);
                const uint32_t lc = (uint32_t)rm[-(int)S - 1] + rm[-1] + rm[S - 1];
                const uint32_t rc = (uint32_t)rm[4 - (int)S] + rm[4] + rm[S + 4];
                const uint32_t q = life_next4_bytes(u, m, d, lc, rc) & valid;       // alive next, 0/1 per byte
                const uint32_t mv = m & valid;
                const uint32_t base = y * side + x0;
                uint32_t sv = 0;
                if (vec) sv = *reinterpret_cast<const uint32_t *>(stab + base);
                else for (uint32_t k = 0; k < nx; ++k) sv |= (uint32_t)(uint8_t)stab[base + k] << (8 * k);
                const uint32_t out = stable_update4_rule(rule, sv, (q & mv) * 255u, (q & ~mv) * 255u, spawn4, max4, min4,
                                                         empty4) & valid;
                acc = __dp4a((int)out, 0x01010101, acc);
                pop += __popc(q);
                // next world: the cells, and the halo images of the ones on an edge
                const uint32_t last = (q >> (8 * (nx - 1))) & 1u;
                for (int ri = 0; ri < 3; ++ri) {
                    if (ri == 1 && y != 0) continue;
                    if (ri == 2 && y != side - 1) continue;
                    uint8_t *row = nxt + (ri == 0 ? y + 1 : (ri == 1 ? side + 1 : 0)) * S;
                    if (nx == 4) *reinterpret_cast<uint32_t *>(row + 4 + x0) = q;
                    else for (uint32_t k = 0; k < nx; ++k) row[4 + x0 + k] = (uint8_t)((q >> (8 * k)) & 1u);
                    if (x0 == 0) row[4 + side] = (uint8_t)(q & 1u);
                    if (x0 + nx == side) row[3] = (uint8_t)last;
                }
                if (vec) *reinterpret_cast<uint32_t *>(stab + base) = out;
                else for (uint32_t k = 0; k < nx; ++k) stab[base + k] = (int8_t)(out >> (8 * k));
            }
        acc = __reduce_add_sync(0xffffffffu, acc);
        pop = __reduce_add_sync(0xffffffffu, pop);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&red[0], acc);
            atomicAdd(reinterpret_cast<unsigned *>(&red[1]), pop);
        }
        __syncthreads();
        const unsigned long long t_step = threadIdx.x == 0 ? globaltimer_ns() : 0ull;
        // the observation: shared memory -> the caller's pinned mirror (16-byte posted writes), fenced by the writers
        if (obs_mirror != nullptr) {
            const uint32_t n16 = size >> 4;
            const uint32_t copiers = n16 < copy_threads ? (n16 ? n16 : 1u) : copy_threads;
            if (threadIdx.x < copiers) {
                for (uint32_t i = threadIdx.x; i < n16; i += copiers)
                    reinterpret_cast<uint4 *>(obs_mirror)[i] = reinterpret_cast<const uint4 *>(stab)[i];
                for (uint32_t i = (n16 << 4) + threadIdx.x; i < size; i += copiers) obs_mirror[i] = stab[i];
                __threadfence_system();
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            // diagnostics first (result[5..6]: ns from "command seen" to "step computed" / to "mirror fenced"), then
            // one 16-byte store: reward, alive and seq arrive together
            const unsigned long long t_pub = globaltimer_ns();
            *reinterpret_cast<volatile int32_t *>(result + 5) = (int32_t)(t_step - t_seen);
            *reinterpret_cast<volatile int32_t *>(result + 6) = (int32_t)(t_pub - t_seen);
            asm volatile("st.volatile.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(result), "r"(red[0]), "r"(red[1]),
                         "r"((int)done), "r"(0) : "memory");
        }
        uint8_t *tp = cur; cur = nxt; nxt = tp;
    }

    // leave: the device planes get the current state back, then the host is told
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) {
        const uint32_t y = i / W, xw = (i - y * W) * 32;
        const uint8_t *row = cur + (y + 1) * S + 4;
        uint32_t word = 0;
        for (uint32_t j = 0; j < 32 && xw + j < side; ++j) word |= (uint32_t)row[xw + j] << j;
        world[i] = word;
    }
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) stable[i] = stab[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int32_t *>(result + 4) = (int32_t)launch_id;
}

}  // namespace cgl

using namespace cgl;

extern "C" uint32_t cgl_sim_step_max_side(void) { return SIM1_MAX_SIDE; }

extern "C" int cgl_sim_step(const uint32_t *world_in, uint32_t *world_out, int8_t *stable, uint32_t side,
                            int32_t action, int spawn, int stable_max, int dead_rule, int empty, int empty_min,
                            int masked_toggle, int8_t *obs_mirror, int32_t *result, uint32_t seq, cgl_stream_t stream)
{
    CGL_REQUIRE(world_in && world_out && stable && side && world_in != world_out, CGL_E_BADARG,
                "cgl_sim_step: bad argument");
    CGL_REQUIRE(side <= SIM1_MAX_SIDE, CGL_E_BADARG, "cgl_sim_step: side must be <= %u", SIM1_MAX_SIDE);
    const uint32_t size = side * side;
    CGL_REQUIRE(action >= 0 && (uint32_t)action <= size, CGL_E_BADINDEX,
                "cgl_sim_step: action %d outside [0, %u]", action, size);
    CGL_REQUIRE(dead_rule >= CGL_DEAD_ZERO && dead_rule <= CGL_DEAD_SAT && empty >= -128 && empty <= 127 &&
                    empty_min >= -128 && empty_min <= 127,
                CGL_E_BADARG, "cgl_sim_step: dead_rule must be 0..2, empty / empty_min must fit int8");
    static PerDeviceOnce once;
    if (once.first())
        CGL_CUDA(cudaFuncSetAttribute(sim1_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SIM1_MAX_SIDE * SIM1_MAX_SIDE + 16 + 4 * SIM1_MAX_SIDE * ((SIM1_MAX_SIDE + 31) / 32)));
    // four cells per thread, whole rows per pass; all rows in one pass where 1024 threads allow it (side <= 64)
    const uint32_t W = cgl_words_per_row(side), tpr = (side + 3) / 4;
    uint32_t rows_per_pass = 1024 / tpr;
    if (rows_per_pass > side) rows_per_pass = side;
    const unsigned threads = (tpr * rows_per_pass + 31) / 32 * 32;
    const size_t smem = ((size + 15u) & ~15u) + 4 * (size_t)side * W;
    sim1_step_kernel<<<1, threads, smem, as_stream(stream)>>>(
        world_in, world_out, stable, side, W, (uint32_t)action, (int8_t)spawn, (int8_t)stable_max, dead_rule,
        (int8_t)empty, (int8_t)empty_min, masked_toggle, obs_mirror, result, seq, tpr, rows_per_pass);
    CGL_LAUNCH_CHECK();
    return 0;
}

// Struct form for bindings: the per-env constants and pointers are filled once, a step passes one pointer plus the
// two values that change (see include/cgl_b200.h).
extern "C" int cgl_sim_step_ex(const cgl_sim_step_args_t *a, int32_t action, uint32_t seq, cgl_stream_t stream)
{
    CGL_REQUIRE(a, CGL_E_BADARG, "cgl_sim_step_ex: null");
    const bool flip = (a->flip_planes != nullptr) && (*a->flip_planes & 1u);
    const int rc = cgl_sim_step(flip ? a->world_b_dev : a->world_a_dev, flip ? a->world_a_dev : a->world_b_dev,
                                a->stable_dev, a->side, action, a->spawn, a->stable_max, a->dead_rule, a->empty,
                                a->empty_min, a->masked_toggle, a->obs_mirror, a->result, seq, stream);
    if (rc == 0 && a->flip_planes != nullptr) ++*a->flip_planes;
    return rc;
}

extern "C" uint32_t cgl_sim_serve_max_side(void) { return SIM1_SERVE_MAX_SIDE; }

// two halo'd byte planes + the stability plane (see sim1_serve_kernel)
static size_t serve_smem_bytes(uint32_t side)
{
    const size_t S = (side + 5u + 3u) & ~3u;
    const size_t plane = ((side + 2) * S + 15u) & ~(size_t)15u;
    return 2 * plane + (((size_t)side * side + 15u) & ~(size_t)15u);
}

// Launch the resident server for one environment (see sim1_serve_kernel).  The world plane that holds the state
// (world_a if *flip_planes is even, else world_b) and the stability plane are read now and written back when the
// kernel leaves; flip_planes is not advanced.
extern "C" int cgl_sim_serve(const cgl_sim_step_args_t *a, const void *cmd_host, uint32_t last_seq, uint32_t launch_id,
                             uint32_t linger_us, cgl_stream_t stream)
{
    CGL_REQUIRE(a && cmd_host && a->world_a_dev && a->world_b_dev && a->stable_dev && a->result && a->side,
                CGL_E_BADARG, "cgl_sim_serve: bad argument");
    CGL_REQUIRE(a->side <= SIM1_SERVE_MAX_SIDE, CGL_E_BADARG, "cgl_sim_serve: side must be <= %u", SIM1_SERVE_MAX_SIDE);
    CGL_REQUIRE(((uintptr_t)a->result & 15u) == 0 && ((uintptr_t)a->obs_mirror & 15u) == 0 && ((uintptr_t)cmd_host & 7u) == 0,
                CGL_E_BADARG, "cgl_sim_serve: result and obs_mirror must be 16-byte aligned, cmd_host 8-byte aligned");
    CGL_REQUIRE(a->dead_rule >= CGL_DEAD_ZERO && a->dead_rule <= CGL_DEAD_SAT && a->empty >= -128 && a->empty <= 127 &&
                    a->empty_min >= -128 && a->empty_min <= 127 && linger_us > 0 && linger_us <= 100000,
                CGL_E_BADARG, "cgl_sim_serve: dead_rule must be 0..2, empty / empty_min must fit int8, linger 1..100000 us");
    static PerDeviceOnce once;
    if (once.first())
        CGL_CUDA(cudaFuncSetAttribute(sim1_serve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)serve_smem_bytes(SIM1_SERVE_MAX_SIDE)));
    const uint32_t side = a->side;
    const uint32_t W = cgl_words_per_row(side), tpr = (side + 3) / 4;
    uint32_t rows_per_pass = 1024 / tpr;
    if (rows_per_pass > side) rows_per_pass = side;
    const unsigned threads = (tpr * rows_per_pass + 31) / 32 * 32;
    const size_t smem = serve_smem_bytes(side);
    static int copiers_knob = -1;                // CGL_SERVE_COPIERS: threads that write + fence the observation mirror (tuning)
    if (copiers_knob < 0) {
        const char *v = getenv("CGL_SERVE_COPIERS");
        copiers_knob = v ? atoi(v) : 0;
    }
    uint32_t copy_threads = copiers_knob > 0 ? (uint32_t)copiers_knob : SERVE_COPY_THREADS;
    if (copy_threads > threads) copy_threads = threads;
    const bool flip = (a->flip_planes != nullptr) && (*a->flip_planes & 1u);
    sim1_serve_kernel<<<1, threads, smem, as_stream(stream)>>>(
        flip ? a->world_b_dev : a->world_a_dev, a->stable_dev, side, W, (int8_t)a->spawn, (int8_t)a->stable_max,
        a->dead_rule, (int8_t)a->empty, (int8_t)a->empty_min, a->masked_toggle, a->obs_mirror, a->result,
        static_cast<const unsigned long long *>(cmd_host), last_seq, launch_id, 1000ull * linger_us, tpr, rows_per_pass,
        copy_threads);
    CGL_LAUNCH_CHECK();
    return 0;
}
