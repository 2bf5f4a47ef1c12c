// tvc_rollout.cu -- fused SAC-actor rollout (BASELINE config 4): T env steps per launch, the actor
// MLP Linear(10,256)-ReLU-Linear(256,256)-ReLU-Linear(256,4) evaluated inside the step loop.
//
// Replaces the get_action -> env.step loop of scripts/train.py:546-603 for the legacy 2x256 SAC actor
// (shape from scripts/export_tflm.py:85-156, tests/test_agent.py:46-56).
//
// sm_100a design: CTA = 512 threads = 512 envs = four 128-row MMA tiles (thread i owns env i and accumulator
// row i % 128 of tile i / 128; the 16 warps share the resident weights and hide each other's physics latency;
// the tiles alternate between two 256-column TMEM accumulator sets and take turns on one hidden-tile buffer).  bf16 weights are packed once into UMMA "K-major, no swizzle" core-matrix images and copied
// into shared memory with one TMA bulk copy (cp.async.bulk) for the whole rollout.  Per step:
//   obs tile -> smem (bf16)          -> tcgen05.mma M128 N256 K16   (layer 1, accumulators in TMEM)
//   tcgen05.ld -> bias+ReLU -> smem  -> 16 x tcgen05.mma M128 N256 K16 (layer 2)
//   tcgen05.ld -> bias+ReLU -> 256->4 head accumulated in registers (0.8 % of the FLOPs, CUDA cores)
//   Philox eps, tanh -> action -> env_pre / integrate / env_post (same device code as step_kernel).
// One elected thread issues the MMAs; completion is signalled through tcgen05.commit -> mbarrier.
#include "tvc_internal.h"
#include "tvc_umma.cuh"

#include <cuda_bf16.h>
#include <cstring>
#include <string>

using namespace tvc;
using namespace tvc_umma;

namespace {

constexpr int HID = 256;
constexpr int K1 = 16;               // layer-1 K (10 obs padded to one UMMA K step)
constexpr int TM = 128;              // rows (envs) of one MMA tile = TMEM lanes
#ifndef TVC_ROLLOUT_BLOCK
#define TVC_ROLLOUT_BLOCK 512
#endif
constexpr int RB = TVC_ROLLOUT_BLOCK; // threads (envs) per CTA = NT tiles; the warps share the resident weights
constexpr int NSET = 2;              // TMEM accumulator column sets (256 columns each); tile j uses set j % 2
constexpr int NT = RB / TM;
constexpr uint32_t W1_BYTES = HID * K1 * 2;        // 8 KB   image [K1/8][256][8] bf16
constexpr uint32_t W2_BYTES = HID * HID * 2;       // 128 KB image [256/8][256][8] bf16
constexpr uint32_t A1_BYTES = TM * K1 * 2;         // 4 KB   obs tile [K1/8][128][8] bf16 (one per tile)
constexpr uint32_t H1_BYTES = TM * HID * 2;        // 64 KB  hidden tile [256/8][128][8] bf16 (shared by the tiles in turn)
constexpr uint32_t VEC_FLOATS = HID + HID + 4 * HID + 4;   // b1, b2, W3[4][256], b3
constexpr uint32_t VEC_BYTES = ((VEC_FLOATS * 4 + 15) / 16) * 16;

// shared-memory map (dynamic smem, 1024-byte aligned base)
constexpr uint32_t OFF_W2 = 0;
constexpr uint32_t OFF_W1 = OFF_W2 + W2_BYTES;
constexpr uint32_t OFF_A1 = OFF_W1 + W1_BYTES;
constexpr uint32_t OFF_H1 = OFF_A1 + NT * A1_BYTES;     // also the contact-exchange area during physics
constexpr uint32_t OFF_VEC = OFF_H1 + H1_BYTES;
constexpr uint32_t OFF_BAR = OFF_VEC + VEC_BYTES;       // weight barrier, one MMA barrier per tile, tmem base
constexpr uint32_t SMEM_TOTAL = OFF_BAR + 64 + 256;   // barriers + tmem base (64 B), class counts of the in-CTA sort (256 B)
static_assert(NT >= 2 && NT % 2 == 0, "tiles alternate between two TMEM column sets");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

// UMMA instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t IDESC = idesc_bf16(128, 256);

struct RolloutWs {
    uint8_t *img = nullptr;   // [W2 image][W1 image][vec] in the shared-memory layout
};

// fp32 [out,in] torch weights -> bf16 UMMA images + fp32 vectors, in the shared-memory layout
__global__ void pack_actor_kernel(tvc_actor_weights w, uint8_t *img) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    // W2 image: element (n, k) at (k/8)*4096 + n*16 + (k%8)*2
    for (int idx = tid; idx < HID * (HID / 8); idx += nthreads) {
        const int n = idx % HID, c = idx / HID;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] = pack_bf16(w.w2[n * HID + 8 * c + 2 * j], w.w2[n * HID + 8 * c + 2 * j + 1]);
        *reinterpret_cast<uint4 *>(img + OFF_W2 + c * (HID * 16) + n * 16) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    // W1 image (K padded 10 -> 16)
    for (int idx = tid; idx < HID * (K1 / 8); idx += nthreads) {
        const int n = idx % HID, c = idx / HID;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const int k = 8 * c + j; v[j] = k < 10 ? w.w1[n * 10 + k] : 0.0f; }
        *reinterpret_cast<uint4 *>(img + OFF_W1 + c * (HID * 16) + n * 16) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    float *vec = reinterpret_cast<float *>(img + W2_BYTES + W1_BYTES);
    for (int idx = tid; idx < (int)VEC_FLOATS; idx += nthreads) {
        float x;
        if (idx < HID) x = w.b1[idx];
        else if (idx < 2 * HID) x = w.b2[idx - HID];
        else if (idx < 6 * HID) x = w.w3[idx - 2 * HID];
        else x = w.b3[idx - 6 * HID];
        vec[idx] = x;
    }
}

struct RolloutIO {
    float *obs;          // [N,10] in: current observation; out: observation after the last step
    float *reward_sum, *actions_last, *actions_all, *reward_all;
    float *obs_all, *next_obs_all;
    uint8_t *term_all, *trunc_all;
    int T, deterministic;
    int per;             // envs per CTA (<= RB): the batch is spread over all SMs when it is large enough (148 CTAs of 443 envs, not 128 of 512)
    unsigned long long t0;   // lifetime step of the first rollout step (Philox counters)
};

template <bool X, int DIV>
__global__ void __launch_bounds__(RB, 1)
rollout_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevState st, const uint8_t *__restrict__ img,
               const __grid_constant__ RolloutIO io) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = tid / TM, row = tid % TM;          // this thread's MMA tile and accumulator row
    // env held by this thread: fixed at first, re-assigned within the CTA every step (class sort, below)
    const long long cta_base = (long long)blockIdx.x * io.per;
    long long i = cta_base + tid;
    bool live = tid < io.per && i < st.n;
    long long gid = c.env_base + i;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_w = s_base + OFF_BAR;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 + 8 * NSET);
    const float *b1 = reinterpret_cast<const float *>(smem + OFF_VEC);
    const float *b2 = b1 + HID;
    const float *w3 = b2 + HID;
    const float *b3 = w3 + 4 * HID;

    // ---- one-time setup: barriers, TMEM (256 accumulator columns per tile), weights via TMA bulk copy ----
    if (tid == 0) {
        mbar_init(bar_w, 1);
#pragma unroll
        for (int j = 0; j < NSET; j++) mbar_init(s_base + OFF_BAR + 8 + 8 * j, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        mbar_expect_tx(bar_w, W2_BYTES + W1_BYTES + VEC_BYTES);
#pragma unroll
        for (uint32_t off = 0; off < W2_BYTES; off += 32768) bulk_g2s(s_base + OFF_W2 + off, img + off, 32768, bar_w);
        bulk_g2s(s_base + OFF_W1, img + W2_BYTES, W1_BYTES, bar_w);
        bulk_g2s(s_base + OFF_VEC, img + W2_BYTES + W1_BYTES, VEC_BYTES, bar_w);
    }

    Env e;
    float obs[10];
    if (live) {
        load_env(st, X, i, e);
#pragma unroll
        for (int k = 0; k < 10; k++) obs[k] = io.obs[10 * i + k];
    } else {
        memset(&e, 0, sizeof(e));
        e.qw = 1.0f; e.pz = 1.0f;
#pragma unroll
        for (int k = 0; k < 10; k++) obs[k] = 0.0f;
    }
    mbar_wait(bar_w, 0);

    uint32_t mma_phase = 0;
    float rsum = 0.0f, a0 = 0.0f, a1 = 0.0f;
    int done = 0, viol = 0;

    for (int t = 0; t < io.T; t++) {
        if (io.obs_all && live) {   // transition record: the observation the actor sees at step t
            float2 *o2 = reinterpret_cast<float2 *>(io.obs_all + ((long long)t * st.n + i) * 10);
#pragma unroll
            for (int k = 0; k < 5; k++) o2[k] = make_float2(obs[2 * k], obs[2 * k + 1]);
        }
        float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
        // All 16 warps work on every tile's epilogues: warp w may read TMEM lanes 32 (w % 4) .. +31 of ANY column, so for each
        // tile it takes row 32 (w % 4) + lane and the 64 accumulator columns 64 (w / 4) .. +63.  (With one tile's four warps
        // doing its 256-column epilogues while the other twelve wait, the MLP cost 19 of the 63 us per step.)  The head's
        // partial sums of a row's four column groups meet through shared memory at the end; layer 2 of tile j runs on the
        // tensor core while the warps are in the second epilogue of tile j - 1.
        {
            const int q = warp & 3, m = warp >> 2;            // TMEM lane quarter, column group
            const int r = q * 32 + lane;                      // accumulator row handled in every tile
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t bar_set0 = s_base + OFF_BAR + 8, bar_set1 = s_base + OFF_BAR + 16;
            {   // layer 1 operands: every env writes its obs row into its own tile's bf16 operand [2][128][8]
                uint4 c0 = make_uint4(pack_bf16(obs[0], obs[1]), pack_bf16(obs[2], obs[3]), pack_bf16(obs[4], obs[5]), pack_bf16(obs[6], obs[7]));
                uint4 c1 = make_uint4(pack_bf16(obs[8], obs[9]), 0u, 0u, 0u);
                uint8_t *a1p = smem + OFF_A1 + tile * A1_BYTES;
                *reinterpret_cast<uint4 *>(a1p + row * 16) = c0;
                *reinterpret_cast<uint4 *>(a1p + TM * 16 + row * 16) = c1;
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {   // layer 1 of tiles 0 and 1 into column sets 0 and 1
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < NSET; j++) {
                    mma_bf16(tmem_base + j * HID, umma_desc(s_base + OFF_A1 + j * A1_BYTES, TM * 16, 128),
                             umma_desc(s_base + OFF_W1, HID * 16, 128), IDESC, 0u);
                    mma_commit(s_base + OFF_BAR + 8 + 8 * j);
                }
            }
            float po[NT][4];
#pragma unroll
            for (int j = 0; j < NT; j++) { po[j][0] = 0.0f; po[j][1] = 0.0f; po[j][2] = 0.0f; po[j][3] = 0.0f; }
            uint32_t ph0 = mma_phase, ph1 = mma_phase;        // running parities of the two column-set barriers
#pragma unroll
            for (int j = 0; j <= NT; j++) {
                if (j < NT) {
                    const int sj = j % NSET;
                    // layer 1 of tile j has landed in set sj (and layer 2 of tile j - 1 has released the hidden tile, below)
                    if (sj == 0) { mbar_wait(bar_set0, ph0); ph0 ^= 1u; } else { mbar_wait(bar_set1, ph1); ph1 ^= 1u; }
                    tc_fence_after();
                    // epilogue 1 of tile j: bias + ReLU -> bf16 hidden tile [32][128][8], this thread's 64 columns of row r
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        uint32_t v[32];
                        const int cb = m * 64 + hh * 32;
                        tmem_ld32(tq + (uint32_t)(sj * HID + cb), v);
#pragma unroll
                        for (int qq = 0; qq < 4; qq++) {
                            float h[8];
#pragma unroll
                            for (int jj = 0; jj < 8; jj++) h[jj] = fmaxf(__uint_as_float(v[8 * qq + jj]) + b1[cb + 8 * qq + jj], 0.0f);
                            *reinterpret_cast<uint4 *>(smem + OFF_H1 + ((cb >> 3) + qq) * (TM * 16) + r * 16) =
                                make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
                        }
                    }
                    fence_async_smem();
                    tc_fence_before();
                    __syncthreads();
                    if (tid == 0) {   // layer 2 of tile j: 16 K-steps of M128 N256 K16 over the hidden tile
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < HID / 16; kk++)
                            mma_bf16(tmem_base + sj * HID, umma_desc(s_base + OFF_H1 + kk * 2 * (TM * 16), TM * 16, 128),
                                     umma_desc(s_base + OFF_W2 + kk * 2 * (HID * 16), HID * 16, 128), IDESC, kk > 0 ? 1u : 0u);
                        mma_commit(s_base + OFF_BAR + 8 + 8 * sj);
                    }
                }
                if (j > 0) {
                    // epilogue 2 + head of tile j - 1 (its layer 2 was waited for below, before the hidden tile was reused)
                    const int jp = j - 1, sp = jp % NSET;
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        uint32_t v[32];
                        const int cb = m * 64 + hh * 32;
                        tmem_ld32(tq + (uint32_t)(sp * HID + cb), v);
#pragma unroll
                        for (int jj = 0; jj < 32; jj++) {
                            const int col = cb + jj;
                            const float h = fmaxf(__uint_as_float(v[jj]) + b2[col], 0.0f);
                            po[jp][0] = fmaf(h, w3[col], po[jp][0]); po[jp][1] = fmaf(h, w3[HID + col], po[jp][1]);
                            po[jp][2] = fmaf(h, w3[2 * HID + col], po[jp][2]); po[jp][3] = fmaf(h, w3[3 * HID + col], po[jp][3]);
                        }
                    }
                    tc_fence_before();
                }
                if (j < NT) {
                    // layer 2 of tile j must be complete before anything touches the hidden tile again / reads its set
                    const int sj = j % NSET;
                    if (sj == 0) { mbar_wait(bar_set0, ph0); ph0 ^= 1u; } else { mbar_wait(bar_set1, ph1); ph1 ^= 1u; }
                    tc_fence_after();
                }
                __syncthreads();   // set (j - 1) % 2 has been read out by every warp; the hidden tile is free
                if (tid == 0 && j > 0 && j + 1 < NT) {   // layer 1 of tile j + 1 into the set tile j - 1 just left
                    tc_fence_after();
                    mma_bf16(tmem_base + ((j + 1) % NSET) * HID, umma_desc(s_base + OFF_A1 + (j + 1) * A1_BYTES, TM * 16, 128),
                             umma_desc(s_base + OFF_W1, HID * 16, 128), IDESC, 0u);
                    mma_commit(s_base + OFF_BAR + 8 + 8 * ((j + 1) % NSET));
                }
            }
            // the head's partial sums: [column group][tile][row] float4 in the (now idle) hidden tile; the env's own thread adds
            // its four groups in a fixed order
            float4 *s_part = reinterpret_cast<float4 *>(smem + OFF_H1);
#pragma unroll
            for (int j = 0; j < NT; j++) s_part[(m * NT + j) * TM + r] = make_float4(po[j][0], po[j][1], po[j][2], po[j][3]);
            __syncthreads();
#pragma unroll
            for (int mm = 0; mm < 4; mm++) {
                const float4 pp = s_part[(mm * NT + tile) * TM + row];
                o0 += pp.x; o1 += pp.y; o2 += pp.z; o3 += pp.w;
            }
            __syncthreads();   // the hidden tile is written again in the next step
        }
        // every column-set barrier completed 2 * NT / NSET (an even number of) phases this step -> parity unchanged
        // ---- action: tanh(mean + exp(clamp(log_std)) * eps), eps from Philox stream 6 ----
        const float mean0 = o0 + b3[0], mean1 = o1 + b3[1];
        const float ls0 = clampf(o2 + b3[2], -20.0f, 2.0f), ls1 = clampf(o3 + b3[3], -20.0f, 2.0f);
        float e0 = 0.0f, e1 = 0.0f;
        if (!io.deterministic) {
            const unsigned long long tt = io.t0 + (unsigned long long)t;
            const uint4 rr = philox(c.seed_lo, c.seed_hi, gid, ST_ACTOR, (unsigned)tt, (unsigned)(tt >> 32));
            box_muller(rr.x, rr.y, e0, e1);
        }
        a0 = tanhf(mean0 + expf(ls0) * e0);
        a1 = tanhf(mean1 + expf(ls1) * e1);

#ifndef TVC_ROLLOUT_NO_SORT
        // ---- class sort within the CTA: the envs change threads so that warps hold envs of one work class ----
        // One thread keeps its accumulator row, not its env: a row of an MMA tile is just a slot, so after the action is
        // known the 512 envs are stably sorted by class (in contact / may touch / airborne) and every thread takes over the
        // env at its own position -- state, action, reward sum and env index travel through the idle operand + hidden-tile
        // area (35 words x 512).  Without it every warp holds a few in-contact envs and walks the whole solver path;
        // with it three or four of the sixteen warps do, and the step is bounded by their (now uncontended) chain.
        {
            int *s_cnt = reinterpret_cast<int *>(smem + OFF_BAR + 64);               // [3][RB / 32]
            float *s_x = reinterpret_cast<float *>(smem + OFF_A1);                   // [35][RB] words
            static_assert(35u * RB * 4u <= NT * A1_BYTES + H1_BYTES, "env exchange must fit in the operand + hidden-tile area");
            static_assert(3 * (RB / 32) * 4 <= 256, "class counts fit behind the barriers");
            const unsigned full = 0xffffffffu;
            const int cls = live ? class_of(c, X, e.pz, e.qx, e.qy, e.qz, e.qw, e.vz, e.wx, e.wy, e.wz, e.cg_off) : 2;
            const unsigned m0 = __ballot_sync(full, cls == 0), m1 = __ballot_sync(full, cls == 1), m2 = __ballot_sync(full, cls == 2);
            if (lane == 0) { s_cnt[warp] = __popc(m0); s_cnt[RB / 32 + warp] = __popc(m1); s_cnt[2 * (RB / 32) + warp] = __popc(m2); }
            __syncthreads();
            int before = 0, t0 = 0, t1 = 0;
#pragma unroll
            for (int w = 0; w < RB / 32; w++) {
                const int c0 = s_cnt[w], c1 = s_cnt[RB / 32 + w], c2 = s_cnt[2 * (RB / 32) + w];
                t0 += c0; t1 += c1;
                if (w < warp) before += cls == 0 ? c0 : (cls == 1 ? c1 : c2);
            }
            const unsigned mk = cls == 0 ? m0 : (cls == 1 ? m1 : m2);
            const int pos = (cls == 0 ? 0 : (cls == 1 ? t0 : t0 + t1)) + before + __popc(mk & ((1u << lane) - 1u));
            float w_[35] = {e.px, e.py, e.pz, e.ep_ret, e.qx, e.qy, e.qz, e.qw, e.vx, e.vy, e.vz, __int_as_float(e.step),
                            e.wx, e.wy, e.wz, __int_as_float(e.burn), __int_as_float(e.phase), __int_as_float(e.success),
                            __int_as_float(e.has_prev), __int_as_float(e.consec), e.ap0, e.ap1, __int_as_float(e.hist_count),
                            __int_as_float(e.n_clip), __int_as_float(e.n_run), e.mass_scale, e.thrust_scale, e.cg_off,
                            e.wind_x, e.wind_y, __int_as_float(e.episode), a0, a1, rsum, __int_as_float((int)(i - cta_base))};
#pragma unroll
            for (int k = 0; k < 35; k++) s_x[k * RB + pos] = w_[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 35; k++) w_[k] = s_x[k * RB + tid];
            e.px = w_[0]; e.py = w_[1]; e.pz = w_[2]; e.ep_ret = w_[3]; e.qx = w_[4]; e.qy = w_[5]; e.qz = w_[6]; e.qw = w_[7];
            e.vx = w_[8]; e.vy = w_[9]; e.vz = w_[10]; e.step = __float_as_int(w_[11]);
            e.wx = w_[12]; e.wy = w_[13]; e.wz = w_[14]; e.burn = __float_as_int(w_[15]); e.phase = __float_as_int(w_[16]);
            e.success = __float_as_int(w_[17]); e.has_prev = __float_as_int(w_[18]); e.consec = __float_as_int(w_[19]);
            e.ap0 = w_[20]; e.ap1 = w_[21]; e.hist_count = __float_as_int(w_[22]); e.n_clip = __float_as_int(w_[23]);
            e.n_run = __float_as_int(w_[24]); e.mass_scale = w_[25]; e.thrust_scale = w_[26]; e.cg_off = w_[27];
            e.wind_x = w_[28]; e.wind_y = w_[29]; e.episode = __float_as_int(w_[30]);
            a0 = w_[31]; a1 = w_[32]; rsum = w_[33];
            i = cta_base + __float_as_int(w_[34]);
            live = __float_as_int(w_[34]) < io.per && i < st.n;
            gid = c.env_base + i;
            __syncthreads();   // the exchange area is the MLP's operand / hidden-tile area again from the next step on
        }
#endif
        // ---- env step (same device code as the step kernel) ----
        BodyP P;
        Forces f;
        if (live) env_pre<X>(c, st, i, e, a0, a1, P, f);
        else {
            P = body_params(c, false, 1.0f, 0.0f, 1.0f);
            f.Fx = f.Fy = f.Fz = f.Tx = f.Ty = f.Tz = f.a0 = f.a1 = f.fl0 = f.fl1 = f.fl2 = f.arm = 0.0f;
        }
        // every thread solves its own env's contacts inline (same device code as the step kernel)
#ifdef TVC_PHASE_PROF2
        Ph2 ph2s = {0u, 0u, 0u};
        if (live) integrate_thread<false, false>(c, P, e, f, &ph2s);
#elif defined(TVC_SOLVE_STATS)
        unsigned ss_mask = 0u;
        if (live) integrate_thread<false, false>(c, P, e, f, ss_mask);
#else
        if (live) integrate_thread<false, false>(c, P, e, f);
#endif
        done = 0; viol = 0;
        int ev_len = 0, ev_succ = 0, ev_reason = 0, ev_trunc = 0;
        float ev_ret = 0.0f, ev_alt = 0.0f, ev_tilt = 0.0f, ev_fuel = 0.0f;
        if (live) {
            StepResult r;
            env_post<X, DIV>(c, st, i, gid, e, f.a0, f.a1, r);
            rsum += r.reward;
            if (io.actions_all) reinterpret_cast<float2 *>(io.actions_all)[(long long)t * st.n + i] = make_float2(a0, a1);
            if (io.reward_all) io.reward_all[(long long)t * st.n + i] = r.reward;
            viol = r.viol;
            done = r.terminated | r.truncated;
            if (io.term_all) io.term_all[(long long)t * st.n + i] = (uint8_t)r.terminated;
            if (io.trunc_all) io.trunc_all[(long long)t * st.n + i] = (uint8_t)r.truncated;
            if (io.next_obs_all) {   // successor observation before any autoreset (the terminal one when done)
                float2 *o2 = reinterpret_cast<float2 *>(io.next_obs_all + ((long long)t * st.n + i) * 10);
#pragma unroll
                for (int k = 0; k < 5; k++) o2[k] = make_float2(r.obs[2 * k], r.obs[2 * k + 1]);
            }
            if (done) {
                ev_len = e.step; ev_succ = e.success; ev_reason = r.reason; ev_trunc = r.truncated;
                ev_ret = e.ep_ret; ev_alt = r.alt; ev_tilt = r.tilt; ev_fuel = r.fuel;
                if (c.autoreset) {
                    reset_env(c, X, gid, e, false);
                    build_obs(c, X, gid, e, 0, r.obs);
                }
            }
#pragma unroll
            for (int k = 0; k < 10; k++) obs[k] = r.obs[k];
        }
        // ---- episode statistics (same scheme as the step kernels; the contact area is idle now) ----
        __syncthreads();
        const int any_ev = __syncthreads_or(done | viol);
        if (any_ev) {
            double *s_stat = reinterpret_cast<double *>(smem + OFF_H1 + 32768);   // [RB/32][16] doubles, clear of the contact area's head
            const unsigned full = 0xffffffffu;
            int n_ep = __reduce_add_sync(full, done), n_len = __reduce_add_sync(full, ev_len);
            int n_succ = __reduce_add_sync(full, done ? ev_succ : 0);
            int n_cr = __reduce_add_sync(full, ev_reason == 2), n_ti = __reduce_add_sync(full, ev_reason == 3);
            int n_al = __reduce_add_sync(full, ev_reason == 4), n_ra = __reduce_add_sync(full, ev_reason == 5);
            int n_tr = __reduce_add_sync(full, ev_trunc), n_vi = __reduce_add_sync(full, viol);
            double d_ret = ev_ret, d_ret2 = (double)ev_ret * (double)ev_ret, d_alt = ev_alt, d_tilt = ev_tilt, d_fuel = ev_fuel;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                d_ret += __shfl_xor_sync(full, d_ret, o); d_ret2 += __shfl_xor_sync(full, d_ret2, o);
                d_alt += __shfl_xor_sync(full, d_alt, o); d_tilt += __shfl_xor_sync(full, d_tilt, o);
                d_fuel += __shfl_xor_sync(full, d_fuel, o);
            }
            if (lane == 0) {
                double *s = s_stat + warp * TVC_NSTAT;
                s[0] = n_ep; s[1] = d_ret; s[2] = d_ret2; s[3] = n_len; s[4] = n_succ; s[5] = n_cr; s[6] = n_ti;
                s[7] = n_al; s[8] = n_ra; s[9] = n_tr; s[10] = n_vi; s[11] = d_alt; s[12] = d_tilt; s[13] = d_fuel;
                s[14] = 0.0; s[15] = 0.0;
            }
            __syncthreads();
            if (tid < TVC_NSTAT) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < RB / 32; w++) s += s_stat[w * TVC_NSTAT + tid];
                if (s != 0.0) st.partial[(long long)blockIdx.x * TVC_NSTAT + tid] += s;
            }
            __syncthreads();
        }
    }

    if (live) {
        store_env(st, X, i, e);
#pragma unroll
        for (int k = 0; k < 10; k++) io.obs[10 * i + k] = obs[k];
        if (io.reward_sum) io.reward_sum[i] = rsum;
        if (io.actions_last) reinterpret_cast<float2 *>(io.actions_last)[i] = make_float2(a0, a1);
    }
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&st.counter[CTR_STEPS], (unsigned)io.T);   // steps since the statistics were reset
    // ---- teardown: release TMEM ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
    }
}

}  // namespace

void tvc_rollout_free(tvc_handle *h) {
    if (!h || !h->rollout_ws) return;
    RolloutWs *ws = static_cast<RolloutWs *>(h->rollout_ws);
    cudaFree(ws->img);
    delete ws;
    h->rollout_ws = nullptr;
}

extern "C" int tvc_rollout(tvc_handle *h, const tvc_actor_weights *w, int32_t T, const tvc_rollout_io *u, tvc_stream stream) {
    if (!h) { tvc_set_err("handle is NULL"); return TVC_E_BADARG; }
    int prev_dev = -1;
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{-1};
    if (cudaGetDevice(&prev_dev) == cudaSuccess && prev_dev != h->device && cudaSetDevice(h->device) == cudaSuccess) restore.d = prev_dev;
    if (!w || !w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->w3 || !w->b3) { tvc_set_err("tvc_rollout: NULL weight pointer"); return TVC_E_BADARG; }
    if (!u || !u->obs) { tvc_set_err("tvc_rollout: io.obs (current observation, [N,10]) must be non-NULL"); return TVC_E_BADARG; }
    if (T < 1 || T > 65536) { tvc_set_err("tvc_rollout: T out of range [1,65536]"); return TVC_E_BADARG; }
    if (!(h->cur.quirks & TVC_Q_FROZEN_FORCES)) { tvc_set_err("tvc_rollout: built for TVC_Q_FROZEN_FORCES (quirk Q3) only"); return TVC_E_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    if (!h->rollout_ws) {
        RolloutWs *ws = new (std::nothrow) RolloutWs();
        if (!ws) { tvc_set_err("out of host memory"); return TVC_E_NOMEM; }
        cudaError_t e = cudaMalloc((void **)&ws->img, W2_BYTES + W1_BYTES + VEC_BYTES);
        if (e != cudaSuccess) { delete ws; tvc_set_err(cudaGetErrorString(e)); return TVC_E_CUDA; }
        h->rollout_ws = ws;
    }
    RolloutWs *ws = static_cast<RolloutWs *>(h->rollout_ws);
    pack_actor_kernel<<<64, 128, 0, s>>>(*w, ws->img);
    RolloutIO io;
    io.obs = u->obs; io.reward_sum = u->reward_sum; io.actions_last = u->actions_last; io.actions_all = u->actions_all;
    io.reward_all = u->reward_all; io.T = T; io.deterministic = u->deterministic;
    io.obs_all = u->obs_all; io.next_obs_all = u->next_obs_all; io.term_all = u->terminated_all; io.trunc_all = u->truncated_all;
    io.t0 = (unsigned long long)h->lifetime_steps;
    // one CTA per SM (222 KB of shared memory): a batch of more than 256 envs per SM is cut into a whole number of waves over
    // every SM (65,536 envs: 148 CTAs of 443 envs instead of 128 CTAs of 512 on 148 SMs; 262,144: 592 of 443 instead of 512 of
    // 512 = three and a half waves); smaller batches keep full CTAs.  One statistics row per CTA.
    int ctas = (int)((h->n + RB - 1) / RB);
    if (h->n > (int64_t)h->num_sms * 256) ctas = (ctas + h->num_sms - 1) / h->num_sms * h->num_sms;
    io.per = (int)((h->n + ctas - 1) / ctas);
    ctas = (int)((h->n + io.per - 1) / io.per);
    const bool X = h->cur.contract == TVC_CONTRACT_X;
    const int dv = h->cur.diversity_mode;
    cudaError_t e = cudaSuccess;
#define GO(XX, DD)                                                                                                   \
    do {                                                                                                             \
        e = cudaFuncSetAttribute(rollout_kernel<XX, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL); \
        if (e == cudaSuccess) rollout_kernel<XX, DD><<<ctas, RB, SMEM_TOTAL, s>>>(h->dc, h->st, ws->img, io); \
    } while (0)
    if (X) { if (dv == 0) GO(true, 0); else if (dv == 1) GO(true, 1); else GO(true, 2); }
    else   { if (dv == 0) GO(false, 0); else if (dv == 1) GO(false, 1); else GO(false, 2); }
#undef GO
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { tvc_set_err(std::string("rollout_kernel: ") + cudaGetErrorString(e)); return TVC_E_CUDA; }
    h->lifetime_steps += T;
    h->order_valid = false;   // the rollout moved the envs behind the step path's sorted sequence
    return TVC_OK;
}
