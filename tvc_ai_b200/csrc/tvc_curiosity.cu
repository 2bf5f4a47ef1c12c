// tvc_curiosity.cu -- row S14: the intrinsic-curiosity term of the env step for the whole batch on the 5th-gen tensor cores.
//
// Reference: env/enhanced_rocket_tvc_env.py:226-269 (CuriosityModule: forward model Linear(10,256)-ReLU-Linear(256,256)-ReLU-
// Linear(256,8) over [state8, action2], never trained -- quirk Q19 -- intrinsic = 0.01 * MSE(prediction, next state8)) and
// :494-506 (added to the reward AFTER the [-1000, 200] clip -- quirk Q14 --, only when state_history is non-empty; the history
// is cleared by reset(), :399-401).  The reference evaluates it on a batch of ONE per env step on the CPU.
//
// sm_100a design: a persistent grid, one CTA per SM, 128 envs per tile.  The bf16 weights live in shared memory for the whole
// launch as UMMA "K-major, no swizzle" images (W2 128 KB + W1 8 KB + W3 8 KB, one TMA bulk copy each); per tile
//   [state8, action2] row -> bf16 operand   -> tcgen05.mma M128 N256 K16        (layer 1, accumulators in TMEM)
//   tcgen05.ld -> bias + ReLU -> bf16 hidden tile -> 16 x tcgen05.mma M128 N256 K16 (layer 2)
//   tcgen05.ld -> bias + ReLU -> bf16 hidden tile -> 16 x tcgen05.mma M128 N16  K16 (layer 3: 8 outputs padded to N = 16)
//   tcgen05.ld (16 columns) -> + b3 -> 0.01 * mean((pred - next8)^2) -> reward, history update
// All 16 warps share the two wide epilogues (warp w reads TMEM lanes 32 (w % 4) .. +31, columns 64 (w / 4) .. +63); the rows'
// global loads for tile j + 1 are issued (branch-free) before tile j's MMAs are waited for, by warps 0-3 for the operand and by
// warps 4-7 for what the narrow epilogue needs.  Everything the vector env did around the
// forward model in ~8 elementwise torch kernels (concat, clamp, where, MSE, reward add, history update) is in the epilogue.
//
// Numerics: operands are bf16, accumulation fp32 -- the intrinsic term agrees with the fp32 torch module to ~3e-3 relative
// and, being ~1e-3 of a reward of 10-100, leaves the reward within 1e-7 relative (north_star bar: 1e-5).  The single-env facade
// (EnhancedRocketTVCEnv) keeps the fp32 torch module, which reproduces the reference's values to 2e-6.
#include "tvc_internal.h"
#include "tvc_umma.cuh"

#include <new>
#include <string>

using namespace tvc_umma;

namespace {

constexpr int HID = 256;
constexpr int K1 = 16;                 // layer-1 K: 8 state + 2 action inputs padded to one UMMA K step
constexpr int TM = 128;                // rows (envs) per tile = TMEM lanes
constexpr int NOUT = 16;               // layer-3 N: 8 outputs padded to the smallest UMMA N for M = 128
constexpr int NTH = 512;
constexpr uint32_t W1_BYTES = HID * K1 * 2;       // 8 KB    [K1/8][256][8] bf16
constexpr uint32_t W2_BYTES = HID * HID * 2;      // 128 KB  [256/8][256][8] bf16
constexpr uint32_t W3_BYTES = NOUT * HID * 2;     // 8 KB    [256/8][16][8] bf16
constexpr uint32_t A1_BYTES = TM * K1 * 2;        // 4 KB    [K1/8][128][8] bf16
constexpr uint32_t H_BYTES = TM * HID * 2;        // 64 KB   [256/8][128][8] bf16 (hidden 1, then hidden 2)
constexpr uint32_t VEC_FLOATS = HID + HID + 8;    // b1, b2, b3
constexpr uint32_t VEC_BYTES = ((VEC_FLOATS * 4 + 1023) / 1024) * 1024;
constexpr uint32_t OFF_W2 = 0;
constexpr uint32_t OFF_W1 = OFF_W2 + W2_BYTES;
constexpr uint32_t OFF_W3 = OFF_W1 + W1_BYTES;
constexpr uint32_t OFF_VEC = OFF_W3 + W3_BYTES;
constexpr uint32_t IMG_BYTES = OFF_VEC + VEC_BYTES;      // what pack_forward_kernel writes and the TMA copies bring in
constexpr uint32_t OFF_A1 = IMG_BYTES;
constexpr uint32_t OFF_H = OFF_A1 + A1_BYTES;
constexpr uint32_t OFF_BAR = OFF_H + H_BYTES;            // weight barrier, three MMA barriers, tmem base
constexpr uint32_t SMEM_TOTAL = OFF_BAR + 64;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
static_assert(OFF_A1 % 1024 == 0 && OFF_H % 1024 == 0, "operand tiles stay 1 KB aligned");
constexpr uint32_t IDESC_WIDE = idesc_bf16(128, 256), IDESC_OUT = idesc_bf16(128, NOUT);
constexpr uint32_t TMEM_COLS = 512;                      // 256 (layers 1 / 2) + 16 (layer 3), power of two
constexpr uint32_t COL_L2 = 256;                         // layer-2 accumulators; layer 3 reuses the first 16 of them

struct CuriosityWs { uint8_t *img = nullptr; bool packed = false; };

// fp32 [out,in] torch weights -> bf16 UMMA images + fp32 bias vectors, in the shared-memory layout
__global__ void pack_forward_kernel(tvc_forward_model w, uint8_t *img) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int idx = tid; idx < HID * (HID / 8); idx += nth) {          // W2: element (n, k) at (k/8) * 4096 + n * 16 + (k%8) * 2
        const int n = idx % HID, c = idx / HID;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] = pack_bf16(w.w2[n * HID + 8 * c + 2 * j], w.w2[n * HID + 8 * c + 2 * j + 1]);
        *reinterpret_cast<uint4 *>(img + OFF_W2 + c * (HID * 16) + n * 16) = make_uint4(p[0], p[1], p[2], p[3]);
    }
    for (int idx = tid; idx < HID * (K1 / 8); idx += nth) {           // W1, K padded 10 -> 16
        const int n = idx % HID, c = idx / HID;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const int k = 8 * c + j; v[j] = k < 10 ? w.w1[n * 10 + k] : 0.0f; }
        *reinterpret_cast<uint4 *>(img + OFF_W1 + c * (HID * 16) + n * 16) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    for (int idx = tid; idx < NOUT * (HID / 8); idx += nth) {         // W3, rows padded 8 -> 16
        const int n = idx % NOUT, c = idx / NOUT;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = n < 8 ? w.w3[n * HID + 8 * c + j] : 0.0f;
        *reinterpret_cast<uint4 *>(img + OFF_W3 + c * (NOUT * 16) + n * 16) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    float *vec = reinterpret_cast<float *>(img + OFF_VEC);
    for (int idx = tid; idx < (int)VEC_FLOATS; idx += nth)
        vec[idx] = idx < HID ? w.b1[idx] : (idx < 2 * HID ? w.b2[idx - HID] : w.b3[idx - 2 * HID]);
}

struct CurIO {
    const float *actions, *obs, *final_obs;
    const uint8_t *term, *trunc;
    float *prev_state;
    uint8_t *has_prev;
    const float *reward_in;
    float *reward_out, *intrinsic;
    int clip_sum;
    long long n;
};

// What one row (env) of a tile needs from global memory, split over two threads so that neither carries all of it through
// the wide epilogues: thread `row` of warps 0-3 feeds the layer-1 operand, thread `row` of warps 4-7 owns the narrow epilogue.
struct RowA { float s[8], a0, a1; };
struct RowB { float ob[8], nx[8], rew; int has_prev, done; };

__device__ __forceinline__ void load_row_a(const CurIO &io, long long env, RowA &r) {
    if (env < io.n) {
        const float4 p0 = reinterpret_cast<const float4 *>(io.prev_state + 8 * env)[0];
        const float4 p1 = reinterpret_cast<const float4 *>(io.prev_state + 8 * env)[1];
        r.s[0] = p0.x; r.s[1] = p0.y; r.s[2] = p0.z; r.s[3] = p0.w; r.s[4] = p1.x; r.s[5] = p1.y; r.s[6] = p1.z; r.s[7] = p1.w;
        const float2 a = reinterpret_cast<const float2 *>(io.actions)[env];
        r.a0 = fminf(fmaxf(a.x, -1.0f), 1.0f); r.a1 = fminf(fmaxf(a.y, -1.0f), 1.0f);      // ref:470 (the env clips the action first)
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) r.s[k] = 0.0f;
        r.a0 = r.a1 = 0.0f;
    }
}

__device__ __forceinline__ void load_row_b(const CurIO &io, long long env, RowB &r) {
    if (env < io.n) {
        r.has_prev = io.has_prev[env];
        r.done = (io.term ? io.term[env] : 0) | (io.trunc ? io.trunc[env] : 0);
        const float2 *o2 = reinterpret_cast<const float2 *>(io.obs + 10 * env);
#pragma unroll
        for (int k = 0; k < 4; k++) { const float2 v = o2[k]; r.ob[2 * k] = v.x; r.ob[2 * k + 1] = v.y; }
        // same-step autoreset: the step's own successor state is the terminal observation.  Loaded unconditionally and
        // selected afterwards (final_obs aliases obs when the caller has none): a load behind a branch on `done` would make
        // the whole prefetch wait for the flag's round trip
        const float2 *f2 = reinterpret_cast<const float2 *>(io.final_obs + 10 * env);
#pragma unroll
        for (int k = 0; k < 4; k++) { const float2 v = f2[k]; r.nx[2 * k] = v.x; r.nx[2 * k + 1] = v.y; }
        r.rew = io.reward_in ? io.reward_in[env] : 0.0f;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) { r.nx[k] = 0.0f; r.ob[k] = 0.0f; }
        r.rew = 0.0f; r.has_prev = 0; r.done = 0;
    }
}

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(NTH, 1)
curiosity_kernel(const uint8_t *__restrict__ img, const __grid_constant__ CurIO io, int ntiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_w = s_base + OFF_BAR, bar1 = bar_w + 8, bar2 = bar_w + 16, bar3 = bar_w + 24;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 32);
    const float *b1 = reinterpret_cast<const float *>(smem + OFF_VEC), *b2 = b1 + HID, *b3 = b2 + HID;

    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar1, 1); mbar_init(bar2, 1); mbar_init(bar3, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {   // the weights: resident for the whole launch
        mbar_expect_tx(bar_w, IMG_BYTES);
#pragma unroll
        for (uint32_t off = 0; off < W2_BYTES; off += 32768) bulk_g2s(s_base + OFF_W2 + off, img + OFF_W2 + off, 32768, bar_w);
        bulk_g2s(s_base + OFF_W1, img + OFF_W1, W1_BYTES, bar_w);
        bulk_g2s(s_base + OFF_W3, img + OFF_W3, W3_BYTES, bar_w);
        bulk_g2s(s_base + OFF_VEC, img + OFF_VEC, VEC_BYTES, bar_w);
    }
    const int q = warp & 3, m = warp >> 2;               // TMEM lane quarter, 64-column group of the wide epilogues
    const int r = q * 32 + lane;                         // accumulator row this thread serves in the wide epilogues
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool feeder = tid < TM;                        // warps 0-3: row = tid, writes the layer-1 operand
    const bool closer = tid >= TM && tid < 2 * TM;       // warps 4-7: row = tid - 128 (TMEM lane quarter = warp - 4), narrow epilogue
    const int crow = tid - TM;

    // One wide epilogue pass: this thread's 32 accumulator columns `col0 .. col0 + 31` of row r -> bias + ReLU -> bf16 -> the
    // hidden tile's K chunks col0 / 8 .. + 3 (thread-contiguous 16-byte stores = the UMMA layout, conflict-free).
    auto wide_pass = [&](uint32_t tcol, int col0, const float *bias) {
        uint32_t v[32];
        tmem_ld32(tq + tcol, v);
#pragma unroll
        for (int qq = 0; qq < 4; qq++) {
            float h[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) h[jj] = fmaxf(__uint_as_float(v[8 * qq + jj]) + bias[col0 + 8 * qq + jj], 0.0f);
            *reinterpret_cast<uint4 *>(smem + OFF_H + ((col0 >> 3) + qq) * (TM * 16) + r * 16) =
                make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
        }
    };
    // eight K steps (hidden columns 128 half .. + 127) of a 256-deep layer: A = the hidden tile, B = a weight image of `rows` rows
    auto mma_half = [&](uint32_t d_tmem, uint32_t w_off, uint32_t rows, uint32_t idesc, int half) {
#pragma unroll
        for (int kk = 8 * half; kk < 8 * half + 8; kk++)
            mma_bf16(d_tmem, umma_desc(s_base + OFF_H + kk * 2 * (TM * 16), TM * 16, 128),
                     umma_desc(s_base + w_off + kk * 2 * (rows * 16), rows * 16, 128), idesc, kk > 0 ? 1u : 0u);
    };
    auto write_operand = [&](const RowA &x) {   // layer-1 operand [2][128][8] bf16: chunk 0 = the eight state inputs, chunk 1 = the two actions + zero padding
        *reinterpret_cast<uint4 *>(smem + OFF_A1 + tid * 16) =
            make_uint4(pack_bf16(x.s[0], x.s[1]), pack_bf16(x.s[2], x.s[3]), pack_bf16(x.s[4], x.s[5]), pack_bf16(x.s[6], x.s[7]));
        *reinterpret_cast<uint4 *>(smem + OFF_A1 + TM * 16 + tid * 16) = make_uint4(pack_bf16(x.a0, x.a1), 0u, 0u, 0u);
    };

    // TMEM columns: [0, 256) layer-1 accumulators, [256, 512) layer-2 accumulators, [256, 272) layer-3 accumulators (written
    // once the first pass of the second epilogue has read them).  The wide epilogues run in two passes of 128 hidden columns
    // (32 per thread); the next layer's first eight K steps are issued after pass 0 and run on the tensor pipe under pass 1,
    // and layer 1 of the NEXT tile is issued as soon as this tile's first epilogue has left its accumulators.
    RowA ca;
    RowB cb_;
    if (feeder) { load_row_a(io, (long long)blockIdx.x * TM + tid, ca); write_operand(ca); }
    if (closer) load_row_b(io, (long long)blockIdx.x * TM + crow, cb_);
    mbar_wait(bar_w, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        mma_bf16(tmem_base, umma_desc(s_base + OFF_A1, TM * 16, 128), umma_desc(s_base + OFF_W1, HID * 16, 128), IDESC_WIDE, 0u);
        mma_commit(bar1);
    }
    const int c32 = m * 32;                              // this thread's 32 columns within a 128-column pass
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ph ^= 1u) {
        RowA na;
        RowB nb;
        const bool more = tile + (int)gridDim.x < ntiles;
        if (more) {     // the next tile's rows: their round trip runs under this tile's first epilogue / its three layers
            if (feeder) load_row_a(io, (long long)(tile + gridDim.x) * TM + tid, na);
            if (closer) load_row_b(io, (long long)(tile + gridDim.x) * TM + crow, nb);
        }
        mbar_wait(bar1, ph);                             // layer 1 of this tile (issued during the previous tile)
        if (tile != (int)blockIdx.x) mbar_wait(bar3, ph ^ 1u);   // layer 3 of the previous tile has left the hidden tile
        tc_fence_after();
        // ---- epilogue 1 (layer-1 accumulators -> hidden 1), layer 2 issued in halves behind it ----
        wide_pass(0u + c32, c32, b1);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); mma_half(tmem_base + COL_L2, OFF_W2, HID, IDESC_WIDE, 0); }
        wide_pass(128u + c32, 128 + c32, b1);
        if (feeder && more) write_operand(na);           // the operand buffer is free: layer 1 of this tile completed above
        fence_async_smem();
        tc_fence_before();
        __syncthreads();                                 // every warp has read the layer-1 accumulators
        if (tid == 0) {
            tc_fence_after();
            mma_half(tmem_base + COL_L2, OFF_W2, HID, IDESC_WIDE, 1);
            mma_commit(bar2);
            if (more) {                                  // layer 1 of the next tile, under this tile's second epilogue
                mma_bf16(tmem_base, umma_desc(s_base + OFF_A1, TM * 16, 128), umma_desc(s_base + OFF_W1, HID * 16, 128), IDESC_WIDE, 0u);
                mma_commit(bar1);
            }
        }
        // ---- epilogue 2 (layer-2 accumulators -> hidden 2), layer 3 issued in halves behind it ----
        mbar_wait(bar2, ph);
        tc_fence_after();
        wide_pass(COL_L2 + c32, c32, b2);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();                                 // columns [256, 384) have been read: layer 3 may write [256, 272)
        if (tid == 0) { tc_fence_after(); mma_half(tmem_base + COL_L2, OFF_W3, NOUT, IDESC_OUT, 0); }
        wide_pass(COL_L2 + 128u + c32, 128 + c32, b2);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) { tc_fence_after(); mma_half(tmem_base + COL_L2, OFF_W3, NOUT, IDESC_OUT, 1); mma_commit(bar3); }
        // ---- the narrow epilogue: one env per thread of warps 4-7 ----
        if (closer) {
            mbar_wait(bar3, ph);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + COL_L2, v);
            const long long env = (long long)tile * TM + crow;
            if (env < io.n) {
                float sq = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; k++) { const float d = (__uint_as_float(v[k]) + b3[k]) - (cb_.done ? cb_.nx[k] : cb_.ob[k]); sq = fmaf(d, d, sq); }
                const float intrinsic = cb_.has_prev ? 0.01f * (sq * 0.125f) : 0.0f;         // ref:268 (MSELoss = mean over the 8 outputs)
                if (io.intrinsic) io.intrinsic[env] = intrinsic;
                if (io.reward_out) {
                    float rw = cb_.rew + intrinsic;                                             // Q14: after the clip (ref:494-502)
                    if (io.clip_sum) rw = fminf(fmaxf(rw, -1000.0f), 200.0f);
                    io.reward_out[env] = rw;
                }
                // history: state_history.append(obs[:8]) (ref:505); a reset clears it (ref:399-401), and under same-step autoreset
                // the row now holds the new episode's first observation, whose step has no predecessor
                reinterpret_cast<float4 *>(io.prev_state + 8 * env)[0] = make_float4(cb_.ob[0], cb_.ob[1], cb_.ob[2], cb_.ob[3]);
                reinterpret_cast<float4 *>(io.prev_state + 8 * env)[1] = make_float4(cb_.ob[4], cb_.ob[5], cb_.ob[6], cb_.ob[7]);
                io.has_prev[env] = cb_.done ? 0 : 1;
            }
            tc_fence_before();
            cb_ = nb;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

}  // namespace

void tvc_curiosity_free(tvc_handle *h) {
    if (!h || !h->curiosity_ws) return;
    CuriosityWs *ws = static_cast<CuriosityWs *>(h->curiosity_ws);
    cudaFree(ws->img);
    delete ws;
    h->curiosity_ws = nullptr;
}

extern "C" int tvc_curiosity(tvc_handle *h, const tvc_forward_model *w, const tvc_curiosity_io *u, tvc_stream stream) {
    if (!h) { tvc_set_err("handle is NULL"); return TVC_E_BADARG; }
    int prev_dev = -1;
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{-1};
    if (cudaGetDevice(&prev_dev) == cudaSuccess && prev_dev != h->device && cudaSetDevice(h->device) == cudaSuccess) restore.d = prev_dev;
    if (w && (!w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->w3 || !w->b3)) { tvc_set_err("tvc_curiosity: NULL weight pointer"); return TVC_E_BADARG; }
    if (!u || !u->actions || !u->obs || !u->prev_state || !u->has_prev) {
        tvc_set_err("tvc_curiosity: actions, obs, prev_state and has_prev must be non-NULL"); return TVC_E_BADARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (!h->curiosity_ws) {
        CuriosityWs *ws = new (std::nothrow) CuriosityWs();
        if (!ws) { tvc_set_err("out of host memory"); return TVC_E_NOMEM; }
        cudaError_t e = cudaMalloc((void **)&ws->img, IMG_BYTES);
        if (e != cudaSuccess) { delete ws; tvc_set_err(cudaGetErrorString(e)); return TVC_E_CUDA; }
        h->curiosity_ws = ws;
    }
    CuriosityWs *ws = static_cast<CuriosityWs *>(h->curiosity_ws);
    if (w) { pack_forward_kernel<<<64, 128, 0, s>>>(*w, ws->img); ws->packed = true; }
    else if (!ws->packed) { tvc_set_err("tvc_curiosity: no weights packed yet (pass the forward model once)"); return TVC_E_STATE; }
    CurIO io;
    io.actions = u->actions; io.obs = u->obs; io.final_obs = u->final_obs ? u->final_obs : u->obs; io.term = u->terminated; io.trunc = u->truncated;
    io.prev_state = u->prev_state; io.has_prev = u->has_prev; io.reward_in = u->reward_in; io.reward_out = u->reward_out;
    io.intrinsic = u->intrinsic; io.clip_sum = u->clip_sum; io.n = h->n;
    const int ntiles = (int)((h->n + TM - 1) / TM);
    const int grid = ntiles < h->num_sms ? ntiles : h->num_sms;
    cudaError_t e = cudaFuncSetAttribute(curiosity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL);
    if (e == cudaSuccess) {
        curiosity_kernel<<<grid, NTH, SMEM_TOTAL, s>>>(ws->img, io, ntiles);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { tvc_set_err(std::string("curiosity_kernel: ") + cudaGetErrorString(e)); return TVC_E_CUDA; }
    return TVC_OK;
}
