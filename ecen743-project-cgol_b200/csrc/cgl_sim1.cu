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
#include "cgl_internal.cuh"

namespace cgl {

constexpr uint32_t SIM1_MAX_SIDE = 320;          // two byte planes: 2 * 320^2 = 204,800 B of shared memory

__global__ void __launch_bounds__(1024)
sim1_step_kernel(const uint32_t *__restrict__ world_in, uint32_t *__restrict__ world_out, int8_t *stable,
                 uint32_t side, uint32_t W, uint32_t action, int8_t spawn, int8_t stable_max, int rule, int8_t empty,
                 int8_t empty_min, int masked, int8_t *obs_mirror, int32_t *result, uint32_t seq)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    const uint32_t size = side * side;
    uint8_t *a = smem_dyn, *b = a + size;
    __shared__ int red[2];
    if (threadIdx.x == 0) { red[0] = 0; red[1] = 0; }
    // world bits -> bytes, the toggle applied on the way (action == size: nothing to toggle)
    for (uint32_t i = threadIdx.x; i < size; i += blockDim.x) {
        const uint32_t y = i / side, x = i - y * side;
        uint8_t v = (world_in[y * W + (x >> 5)] >> (x & 31)) & 1u;
        if (i == action) v ^= 1u;
        a[i] = v;
    }
    __syncthreads();

    // four cells per thread and trip: one 32-bit access to the stability plane and to the mirror
    int acc = 0;
    for (uint32_t base = threadIdx.x * 4; base < size; base += blockDim.x * 4) {
        const uint32_t n = size - base < 4 ? size - base : 4;
        uint32_t sv = 0;
        if (n == 4) sv = *reinterpret_cast<const uint32_t *>(stable + base);
        else for (uint32_t k = 0; k < n; ++k) sv |= (uint32_t)(uint8_t)stable[base + k] << (8 * k);
        uint32_t out = 0;
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t i = base + k;
            const uint32_t y = i / side, x = i - y * side;
            const uint32_t yu = (y == 0 ? side : y) - 1, yd = (y + 1 == side) ? 0 : y + 1;
            const uint32_t xl = (x == 0 ? side : x) - 1, xr = (x + 1 == side) ? 0 : x + 1;
            const uint32_t cnt = a[yu * side + xl] + a[yu * side + x] + a[yu * side + xr] + a[y * side + xl] +
                                 a[y * side + xr] + a[yd * side + xl] + a[yd * side + x] + a[yd * side + xr];
            const uint8_t p = a[i];
            const uint8_t q = (cnt == 3u) || (cnt == 2u && p);
            b[i] = q;
            int8_t s = (int8_t)(sv >> (8 * k));
            // toggle_state: SPAWN (base env: even when toggled to dead, N2; fork: 0 then)
            if (i == action) s = (masked && !p) ? (int8_t)0 : spawn;
            s = stable_update1_rule(rule, s, p != 0, q != 0, spawn, stable_max, empty, empty_min);
            acc += s;
            out |= (uint32_t)(uint8_t)s << (8 * k);
        }
        if (n == 4) {
            *reinterpret_cast<uint32_t *>(stable + base) = out;
            if (obs_mirror != nullptr) *reinterpret_cast<uint32_t *>(obs_mirror + base) = out;
        } else {
            for (uint32_t k = 0; k < n; ++k) {
                stable[base + k] = (int8_t)(out >> (8 * k));
                if (obs_mirror != nullptr) obs_mirror[base + k] = (int8_t)(out >> (8 * k));
            }
        }
    }
    if (obs_mirror != nullptr) __threadfence_system();       // my mirror stores are visible to the host ...
    __syncthreads();                                          // ... before anyone publishes the sequence number

    uint32_t pop = 0;
    for (uint32_t wdx = threadIdx.x; wdx < side * W; wdx += blockDim.x) {
        const uint32_t y = wdx / W, x0 = (wdx - y * W) * 32;
        uint32_t word = 0;
        for (uint32_t j = 0; j < 32 && x0 + j < side; ++j) word |= (uint32_t)b[y * side + x0 + j] << j;
        pop += __popc(word);
        world_out[wdx] = word;
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
                                      2 * SIM1_MAX_SIDE * SIM1_MAX_SIDE));
    // a thread serves four cells per trip; one trip for side <= 64
    unsigned threads = ((size + 3) / 4 + 31) / 32 * 32;
    threads = threads > 1024 ? 1024 : threads;
    sim1_step_kernel<<<1, threads, 2 * size, as_stream(stream)>>>(
        world_in, world_out, stable, side, cgl_words_per_row(side), (uint32_t)action, (int8_t)spawn,
        (int8_t)stable_max, dead_rule, (int8_t)empty, (int8_t)empty_min, masked_toggle, obs_mirror, result, seq);
    CGL_LAUNCH_CHECK();
    return 0;
}
