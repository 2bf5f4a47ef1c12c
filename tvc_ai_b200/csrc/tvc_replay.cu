// tvc_replay.cu -- uniform-sample gather from the on-device replay ring.
//
// The ring itself is caller memory (torch tensors): tvc_rollout writes the transitions of T steps straight into the ring at
// its head (tvc_rollout_io.obs_all / actions_all / reward_all / next_obs_all / terminated_all point into it), so filling the
// buffer costs no copy.  This kernel draws `batch` uniform indices from Philox4x32-10 (key = seed, counter = (sample, draw))
// and gathers the five arrays of a transition into the learner's static batch tensors in ONE launch -- the replacement of
// scripts/train.py:574-584's batch-of-1 dict for agent/multi_algorithm_agent.py:950-1016 (_update_sac).
#include "tvc_internal.h"

#include <string>

using namespace tvc;

namespace {

// one warp per sample: lanes 0-9 move obs, 10-19 next_obs, 20-21 the action, 22 the reward, 23 the done flag
__global__ void __launch_bounds__(256)
replay_gather_kernel(const float *__restrict__ obs, const float *__restrict__ act, const float *__restrict__ rew,
                     const float *__restrict__ nobs, const uint8_t *__restrict__ term, unsigned long long filled, int batch,
                     unsigned seed_lo, unsigned seed_hi, unsigned long long draw, const unsigned long long *__restrict__ ctl, float reward_scale,
                     float *__restrict__ o_obs, float *__restrict__ o_act, float *__restrict__ o_rew, float *__restrict__ o_nobs,
                     float *__restrict__ o_done, long long *__restrict__ o_idx) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (b >= batch) return;
    unsigned long long j = 0;
    if (lane == 0) {
        if (ctl) { filled = ctl[0]; draw += ctl[1]; }   // device-side control words: the launch can be replayed from a CUDA graph
        const uint4 r = philox(seed_lo, seed_hi, (long long)b, 8u /* stream: replay */, (unsigned)draw, (unsigned)(draw >> 32));
        // 64 random bits -> [0, filled): multiply-shift (bias < 2^-40 for any realistic ring size)
        const unsigned long long x = ((unsigned long long)r.x << 32) | r.y;
        j = __umul64hi(x, filled);
    }
    j = __shfl_sync(0xffffffffu, j, 0);
    if (lane < 10) o_obs[(long long)b * 10 + lane] = obs[j * 10 + lane];
    else if (lane < 20) o_nobs[(long long)b * 10 + (lane - 10)] = nobs[j * 10 + (lane - 10)];
    else if (lane < 22) o_act[(long long)b * 2 + (lane - 20)] = act[j * 2 + (lane - 20)];
    else if (lane == 22) o_rew[b] = rew[j] * reward_scale;
    else if (lane == 23) o_done[b] = term[j] ? 1.0f : 0.0f;
    else if (lane == 24 && o_idx) o_idx[b] = (long long)j;
}

}  // namespace

extern "C" int tvc_replay_sample(const tvc_replay_ring *ring, int64_t filled, int32_t batch, uint64_t seed, uint64_t draw,
                                 const uint64_t *ctl_dev, float reward_scale, const tvc_replay_batch *out, int device, tvc_stream stream) {
    if (!ring || !out || !ring->obs || !ring->actions || !ring->reward || !ring->next_obs || !ring->terminated ||
        !out->obs || !out->actions || !out->reward || !out->next_obs || !out->done) {
        tvc_set_err("tvc_replay_sample: NULL pointer"); return TVC_E_BADARG;
    }
    if (filled < 1 || filled > ring->capacity || batch < 1) { tvc_set_err("tvc_replay_sample: filled / batch out of range"); return TVC_E_BADARG; }
    int prev = -1;
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{-1};
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device && cudaSetDevice(device) == cudaSuccess) restore.d = prev;
    const int warps_per_cta = 256 / 32;
    replay_gather_kernel<<<(batch + warps_per_cta - 1) / warps_per_cta, 256, 0, (cudaStream_t)stream>>>(
        ring->obs, ring->actions, ring->reward, ring->next_obs, ring->terminated, (unsigned long long)filled, batch, (unsigned)seed,
        (unsigned)(seed >> 32), (unsigned long long)draw, (const unsigned long long *)ctl_dev, reward_scale, out->obs, out->actions, out->reward, out->next_obs, out->done,
        (long long *)out->indices);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { tvc_set_err(std::string("replay_gather_kernel: ") + cudaGetErrorString(e)); return TVC_E_CUDA; }
    return TVC_OK;
}
