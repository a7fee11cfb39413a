// cgl_rollout.cu -- host-driven rollouts: the env step fed from HOST buffers every step, with the host out of the
// GPU's way.
//
// Reference: the body of the training loop, /root/reference/CGL/main.py:64-72 -- per step the host hands the env an
// action (`toggle_state`, CGL/CGL.py:322-328), steps it (`step`, :247-252; the reference moves both planes over
// PCIe in both directions, :203-208) and reads the reward back (`reward`, :255-256).  Here the state stays on the
// device; what crosses PCIe per step is the action word in and the reward word out of every environment.
//
// A synchronous host loop (copy, launch, synchronise, think, repeat) leaves the GPU idle while the host thinks and
// the host idle while the GPU steps.  The rollout splits the batch into n_groups GROUPS that are stepped
// alternately, each on its own stream: while group g's step runs, the host collects group g+1's rewards, asks the
// policy for its next actions and enqueues its next step.  The data dependence the loop has -- an env's next
// action may depend on its last reward -- is kept per group: the policy is called for a group only after that
// group's previous step has completed.
//
//   * one group step = ONE cudaGraphLaunch of [H2D copy of the group's pinned action buffer -> fused env kernel
//     (-> D2H copy of the observation, if asked for)], instantiated once per (group, replica, plane orientation);
//   * rewards are written by the kernel straight into pinned, host-mapped memory (4 B per env, posted writes);
//   * completion is an event the host polls (cudaEventQuery spin, no blocking synchronisation);
//   * n_replicas > 1 rotates every group over several resident env batches (step s uses replica s % n_replicas) --
//     bench.py uses it so that the working set exceeds the L2 like in the device-timed measurement.
#include <new>
#include <vector>

#include "cgl_internal.cuh"

namespace cgl {
extern int g_pdl_suppress;
}

struct cgl_rollout {
    struct Replica {
        uint32_t *plane[2];
        int8_t *stable;
        int parity;                         // plane[parity] holds the current world
        cudaGraphExec_t exec[2];            // indexed by the parity the step starts from
    };
    struct Group {
        std::vector<Replica> rep;
        int32_t *act_host, *act_dev, *act_dev_alias, *rew_host, *rew_dev_alias, *rew_dev;
        int8_t *obs_host;
        cudaStream_t st;
        cudaEvent_t ev;
        bool busy;
        uint64_t steps_done;
    };
    std::vector<Group> g;
    uint64_t n;                             // envs per group
    uint32_t side, n_replicas;
    int spawn, stable_max, device;
    bool zero_copy;                         // the kernel reads the actions straight from pinned host memory
    uint64_t step;                          // global step counter (all groups stepped `step` times)
};

using namespace cgl;

static int rollout_wait(cgl_rollout::Group &gr)
{
    if (!gr.busy) return 0;
    cudaError_t e;
    while ((e = cudaEventQuery(gr.ev)) == cudaErrorNotReady) {
    }
    if (e != cudaSuccess) {
        set_error("cgl_rollout: group step failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    gr.busy = false;
    return 0;
}

extern "C" int cgl_rollout_destroy(cgl_rollout_t *r)
{
    if (r == nullptr) return 0;
    for (auto &gr : r->g) {
        if (gr.st) cudaStreamSynchronize(gr.st);
        for (auto &rp : gr.rep)
            for (int p = 0; p < 2; ++p)
                if (rp.exec[p]) cudaGraphExecDestroy(rp.exec[p]);
        if (gr.ev) cudaEventDestroy(gr.ev);
        if (gr.st) cudaStreamDestroy(gr.st);
        if (gr.act_host) cudaFreeHost(gr.act_host);
        if (gr.rew_host) cudaFreeHost(gr.rew_host);
        if (gr.act_dev) cudaFree(gr.act_dev);
        if (gr.rew_dev) cudaFree(gr.rew_dev);
    }
    delete r;
    return 0;
}

extern "C" int cgl_rollout_create(cgl_rollout_t **out, uint32_t n_groups, uint32_t n_replicas,
                                  uint32_t *const *world_a_dev, uint32_t *const *world_b_dev,
                                  int8_t *const *stable_dev, uint64_t envs_per_group, uint32_t side, int spawn,
                                  int stable_max, int8_t *const *obs_host, uint32_t flags)
{
    CGL_REQUIRE(out && n_groups >= 1 && n_groups <= 16 && n_replicas >= 1 && n_replicas <= 64 && world_a_dev &&
                    world_b_dev && stable_dev && envs_per_group && side,
                CGL_E_BADARG, "cgl_rollout_create: bad argument");
    cgl_rollout *r = new (std::nothrow) cgl_rollout();
    CGL_REQUIRE(r, CGL_E_NOMEM, "cgl_rollout_create: out of memory");
    r->n = envs_per_group; r->side = side; r->n_replicas = n_replicas; r->spawn = spawn; r->stable_max = stable_max;
    r->step = 0;
    r->zero_copy = (flags & CGL_ROLLOUT_ZERO_COPY_ACTIONS) != 0;
    int rc = 0;
#define RB_CUDA(expr)                                                                                     \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);        \
            cgl_rollout_destroy(r);                                                                       \
            return (int)_e;                                                                               \
        }                                                                                                 \
    } while (0)
    RB_CUDA(cudaGetDevice(&r->device));
    r->g.resize(n_groups);
    const size_t abytes = envs_per_group * sizeof(int32_t);
    for (uint32_t gi = 0; gi < n_groups; ++gi) {
        auto &gr = r->g[gi];
        gr = cgl_rollout::Group();
        gr.busy = false;
        gr.steps_done = 0;
        gr.obs_host = obs_host ? obs_host[gi] : nullptr;
        RB_CUDA(cudaStreamCreateWithFlags(&gr.st, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&gr.ev, cudaEventDisableTiming));
        RB_CUDA(cudaHostAlloc(&gr.act_host, abytes, cudaHostAllocMapped | cudaHostAllocPortable));
        RB_CUDA(cudaHostGetDevicePointer(&gr.act_dev_alias, gr.act_host, 0));
        RB_CUDA(cudaHostAlloc(&gr.rew_host, abytes, cudaHostAllocMapped | cudaHostAllocPortable));
        RB_CUDA(cudaHostGetDevicePointer(&gr.rew_dev_alias, gr.rew_host, 0));
        RB_CUDA(cudaMalloc(&gr.act_dev, abytes));
        // fused sides: one thread per env writes the reward -- straight into the pinned buffer.  Generic sides
        // accumulate it with atomics, which stay on the device; a D2H copy node follows the step.
        const bool direct = cgl_env_step_is_fused(side) != 0;
        if (!direct) RB_CUDA(cudaMalloc(&gr.rew_dev, abytes));
        memset(gr.act_host, 0, abytes);
        memset(gr.rew_host, 0, abytes);
        for (uint64_t e = 0; e < envs_per_group; ++e) gr.act_host[e] = (int32_t)(side * side);     // "do nothing"
        gr.rep.resize(n_replicas);
        for (uint32_t ri = 0; ri < n_replicas; ++ri) {
            auto &rp = gr.rep[ri];
            const size_t k = (size_t)gi * n_replicas + ri;
            rp.plane[0] = world_a_dev[k]; rp.plane[1] = world_b_dev[k]; rp.stable = stable_dev[k];
            rp.parity = 0; rp.exec[0] = rp.exec[1] = nullptr;
            if (!(rp.plane[0] && rp.plane[1] && rp.stable)) {
                set_error("cgl_rollout_create: null plane for group %u replica %u", gi, ri);
                cgl_rollout_destroy(r);
                return CGL_E_BADARG;
            }
            for (int p = 0; p < 2; ++p) {        // one graph per plane orientation: copy -> step (-> obs copy)
                cudaGraph_t graph = nullptr;
                RB_CUDA(cudaStreamBeginCapture(gr.st, cudaStreamCaptureModeThreadLocal));
                cudaError_t e1 = cudaSuccess;
                if (!r->zero_copy) e1 = cudaMemcpyAsync(gr.act_dev, gr.act_host, abytes, cudaMemcpyHostToDevice, gr.st);
                g_pdl_suppress = 1;              // (the step follows a copy node: nothing to overlap with)
                rc = cgl_env_step(rp.plane[p], rp.plane[p ^ 1], rp.stable, envs_per_group, side,
                                  r->zero_copy ? gr.act_dev_alias : gr.act_dev, spawn,
                                  stable_max, direct ? gr.rew_dev_alias : gr.rew_dev, nullptr, nullptr, gr.st);
                g_pdl_suppress = 0;
                cudaError_t e2 = cudaSuccess;
                if (!direct) e2 = cudaMemcpyAsync(gr.rew_host, gr.rew_dev, abytes, cudaMemcpyDeviceToHost, gr.st);
                if (gr.obs_host && e2 == cudaSuccess)
                    e2 = cudaMemcpyAsync(gr.obs_host, rp.stable, envs_per_group * (size_t)side * side,
                                         cudaMemcpyDeviceToHost, gr.st);
                cudaError_t e3 = cudaStreamEndCapture(gr.st, &graph);
                if (rc || e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || graph == nullptr) {
                    if (!rc) set_error("cgl_rollout_create: stream capture failed: %s",
                                       cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
                    if (graph) cudaGraphDestroy(graph);
                    cgl_rollout_destroy(r);
                    return rc ? rc : (int)(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3);
                }
                cudaError_t e4 = cudaGraphInstantiate(&rp.exec[p], graph, 0);
                cudaGraphDestroy(graph);
                RB_CUDA(e4);
            }
        }
    }
#undef RB_CUDA
    *out = r;
    return 0;
}

extern "C" int cgl_rollout_buffers(cgl_rollout_t *r, uint32_t group, int32_t **actions_host, int32_t **reward_host)
{
    CGL_REQUIRE(r && group < r->g.size(), CGL_E_BADARG, "cgl_rollout_buffers: bad argument");
    if (actions_host) *actions_host = r->g[group].act_host;
    if (reward_host) *reward_host = r->g[group].rew_host;
    return 0;
}

extern "C" int cgl_rollout_run(cgl_rollout_t *r, uint64_t steps, cgl_policy_fn policy, void *user)
{
    CGL_REQUIRE(r, CGL_E_BADARG, "cgl_rollout_run: null");
    int rc;
    for (uint64_t s = 0; s < steps; ++s) {
        const uint32_t ri = (uint32_t)(r->step % r->n_replicas);
        for (uint32_t gi = 0; gi < r->g.size(); ++gi) {
            auto &gr = r->g[gi];
            if ((rc = rollout_wait(gr))) return rc;               // this group's previous step has landed ...
            if (policy) policy(user, gi, r->step, gr.rew_host, gr.act_host);      // ... the host decides ...
            auto &rp = gr.rep[ri];
            CGL_CUDA(cudaGraphLaunch(rp.exec[rp.parity], gr.st));                  // ... and the next step is enqueued
            CGL_CUDA(cudaEventRecord(gr.ev, gr.st));
            rp.parity ^= 1;
            gr.busy = true;
            ++gr.steps_done;
        }
        ++r->step;
    }
    for (auto &gr : r->g)
        if ((rc = rollout_wait(gr))) return rc;
    return 0;
}

extern "C" int cgl_rollout_parity(const cgl_rollout_t *r, uint32_t group, uint32_t replica)
{
    if (!r || group >= r->g.size() || replica >= r->n_replicas) return CGL_E_BADARG;
    return r->g[group].rep[replica].parity;
}
