// tvc_abi.cu -- kernels + the C ABI declared in include/tvc_b200.h (libtvc_b200.so).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/tvc_b200.h"
#include "tvc_device.cuh"
#include "tvc_internal.h"

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace tvc;

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------

// Coalesced [rows,10] fp32 store of one warp's observations through shared memory.
__device__ __forceinline__ void warp_store_obs(float *dst, const float *s_warp, long long base, long long n, int lane) {
    long long cnt = n - base;
    if (cnt <= 0) return;
    if (cnt > 32) cnt = 32;
    const int n2 = (int)cnt * 5;   // float2 elements; base*10 floats is 8-byte aligned
    float2 *d2 = reinterpret_cast<float2 *>(dst + base * 10);
    const float2 *s2 = reinterpret_cast<const float2 *>(s_warp);
    for (int k = lane; k < n2; k += 32) d2[k] = s2[k];
}

// ------------------------------------------------------------------------------------------
// step path: ONE class-ordered work sequence over all envs, one warp per 32-env group, solver inline, no CTA barriers
// ------------------------------------------------------------------------------------------
#ifndef TVC_CHUNK
#define TVC_CHUNK 1024   // envs whose class counts are kept together (and whose part of the sequence one CTA writes)
#endif
#define TVC_SUPER 256    // chunks per super-chunk (second level of the counts: the prefix of a chunk is O(nsuper + TVC_SUPER))

// Heuristic class of an env for the coming step (class_of): 0 = its lowest point touches the ground now (measured: such
// envs need the contact solve in 9.4 of the 10 substeps), 1 = it may come within reach during the step, 2 = airborne.
// Only the grouping depends on it (every thread carries the solver), never the results.
//
// The sequence step_kernel_v2 walks -- all class-0 envs, then class 1, then class 2, each in env order -- is the flat array
// st.order.  Its producer is split over the two kernels of a step so that almost nothing of it is on the critical path:
//   * whoever decides an env's class (step_kernel_v2 at the end of the env's step; classify_state_kernel after a reset /
//     set_state / rollout) stores the class byte and adds 1 to the (class, chunk) and (class, super-chunk) counters of the
//     NEXT sequence -- one warp-aggregated integer reduction per distinct (class, chunk) in the warp (deterministic: sums);
//   * close_kernel, the second and last launch of a step, turns that into the sequence: scatter CTA b sums the counters in
//     front of chunk b (exclusive prefix, two levels), ranks its chunk's 1,024 class bytes with ballots and writes the env
//     ids to their global positions.  No sort, no scan pass, no last-CTA tail: ~3 us for 262,144 envs.
// The counters (and the done-list length) are double-buffered by the parity of a sequence number S kept ON THE DEVICE
// (a host-side parity would be baked into a captured CUDA graph).  No kernel reads a word that the same launch writes, so the
// hand-over needs no ticket, fence or last-CTA tail: the step kernel reads S (CTR_SEQ) and leaves a copy (CTR_SEQ2); it adds
// into count buffer (S & 1) ^ 1 and appends to the done list of length CTR_DONE + (S & 1); close_kernel reads the copy, builds
// the sequence from buffer (S & 1) ^ 1, clears buffer S & 1, and one of its threads writes S + 1, zeroes the queue head and
// the other done-list length.
//
// An env whose episode ended gets class 2 for the next sequence (a fresh env starts at z = 1 m); it is re-initialised by the
// reset CTAs of close_kernel.
__device__ __forceinline__ void count_class(const DevState &st, int buf, bool live, long long env, int cls) {
    // lanes of the warp with the same (chunk, class) elect a leader, which adds their number to both counter levels
    const int chunk = (int)(env / TVC_CHUNK);
    const int key = live ? (chunk * 4 + cls) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (live && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        const int cnt = __popc(peers);
        int *cc = st.ccount + (long long)buf * 3 * st.nchunks, *sc = st.scount + buf * 3 * st.nsuper;
        atomicAdd(cc + (long long)cls * st.nchunks + chunk, cnt);
        atomicAdd(sc + cls * st.nsuper + chunk / TVC_SUPER, cnt);
    }
}

// Classes from the state planes (80 B per env): first step, or a reset / set_state / rollout / curriculum change moved the
// envs behind the step path's back.  Writes the class bytes and adds to the counters (both buffers zeroed by the host before).
template <bool X>
__global__ void __launch_bounds__(256)
classify_state_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevState st) {
    const long long env = (long long)blockIdx.x * 256 + threadIdx.x;
    const bool live = env < st.n;
    int cls = 2;
    if (live) {
        const float4 p = st.s0[env], q = st.s1[env], v = st.s2[env], w = st.s3[env];
        cls = class_of(c, X, p.z, q.x, q.y, q.z, q.w, v.z, w.x, w.y, w.z, X ? st.d0[env].z : 0.0f);
        st.cls[env] = (uint8_t)cls;
    }
    const unsigned seq = st.counter[CTR_SEQ];
    if (env == 0) st.counter[CTR_SEQ2] = seq;
    count_class(st, (int)(seq & 1u) ^ 1, live, live ? env : 0, cls);   // like a step: the buffer the following close_kernel reads
}

// One warp writes chunk b's part of the next sequence: the exclusive prefix of the chunk (super-chunks in front, then the
// chunks of its own super-chunk) and the class totals from the counters of buffer `buf`, then 32 rounds of 32 class bytes
// ranked with ballots.  Clears the chunk's counters of the other buffer (the next step adds there).
__device__ __forceinline__ void scatter_chunk(const DevState &st, int b, int buf, int lane) {
    const unsigned full = 0xffffffffu;
    const int nc = st.nchunks, ns = st.nsuper, sup = b / TVC_SUPER;
    const int *cc = st.ccount + (long long)buf * 3 * nc, *sc = st.scount + buf * 3 * ns;
    long long pos[3];
    {
        int pre[3], tot[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int p = 0, t = 0;
            for (int s0 = lane; s0 < ns; s0 += 32) { const int v = __ldcg(sc + k * ns + s0); t += v; if (s0 < sup) p += v; }
            for (int ch = sup * TVC_SUPER + lane; ch < b; ch += 32) p += __ldcg(cc + (long long)k * nc + ch);
            pre[k] = __reduce_add_sync(full, p); tot[k] = __reduce_add_sync(full, t);
        }
        pos[0] = pre[0]; pos[1] = (long long)tot[0] + pre[1]; pos[2] = (long long)tot[0] + tot[1] + pre[2];
        if (b == 0 && lane < 3) st.totals[lane] = lane == 0 ? tot[0] : (lane == 1 ? tot[1] : tot[2]);
    }
    const long long base = (long long)b * TVC_CHUNK;
    static_assert(TVC_CHUNK == 1024, "a lane holds the class bytes of 32 envs of the chunk");
    int cl[32];
#pragma unroll
    for (int j = 0; j < 32; j++) {      // all 32 byte loads of the lane in flight at once (one round trip)
        const long long env = base + j * 32 + lane;
        cl[j] = env < st.n ? (int)__ldcg(st.cls + env) : 3;
    }
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const long long env = base + j * 32 + lane;
        const unsigned m0 = __ballot_sync(full, cl[j] == 0), m1 = __ballot_sync(full, cl[j] == 1), m2 = __ballot_sync(full, cl[j] == 2);
        const unsigned mk = cl[j] == 0 ? m0 : (cl[j] == 1 ? m1 : m2);
        const long long at = (cl[j] == 0 ? pos[0] : (cl[j] == 1 ? pos[1] : pos[2])) + __popc(mk & ((1u << lane) - 1u));
        if (cl[j] < 3) st.order[at] = (int)env;
        pos[0] += __popc(m0); pos[1] += __popc(m1); pos[2] += __popc(m2);
    }
    if (lane < 3) {
        st.ccount[(long long)(buf ^ 1) * 3 * nc + (long long)lane * nc + b] = 0;
        if (b % TVC_SUPER == 0) st.scount[(buf ^ 1) * 3 * ns + lane * ns + sup] = 0;
    }
}

// The second and last launch of a step (also follows classify_state_kernel).  CTAs [0, nreset): same-step autoreset -- one
// env of st.done_list per thread, full warps (ref:381-464 reset + the reset observation over the terminal one; the terminal
// state was stored by the step kernel and every Env field round-trips through store_env / load_env, so this equals resetting
// in place); grid-stride, so any number of episodes may end in one step.  CTAs [nreset, ...): one warp per chunk of the next
// sequence (scatter_chunk).  The last CTA to finish hands the counters to the next step.
template <bool X>
__global__ void __launch_bounds__(128)
close_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevState st, float *obs, int nreset, int count_step) {
    const int lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.wait;" ::: "memory");              // behind step_kernel_v2 / classify_state_kernel
    asm volatile("griddepcontrol.launch_dependents;");              // the next step kernel may be scheduled; it waits for this grid
    const unsigned seq = st.counter[CTR_SEQ2];                      // the sequence number the kernel before this one ran with
    if ((int)blockIdx.x < nreset) {
        unsigned k = blockIdx.x * 128u + threadIdx.x;
        long long i = st.done_list[k < (unsigned)st.n ? k : 0u];   // issued beside the length, not behind it
        const unsigned ndone = st.counter[CTR_DONE + (seq & 1u)];
        while (k < ndone) {
            const long long gid = c.env_base + i;
            Env e;
            load_env(st, X, i, e);
            reset_env(c, X, gid, e, false);
            store_env(st, X, i, e);
            float o[10];
            build_obs(c, X, gid, e, 0, o);
            float2 *o2 = reinterpret_cast<float2 *>(obs + 10 * i);
#pragma unroll
            for (int q = 0; q < 5; q++) o2[q] = make_float2(o[2 * q], o[2 * q + 1]);
            k += (unsigned)nreset * 128u;
            if (k < ndone) i = st.done_list[k];
        }
    } else {
        const int b = (int)((blockIdx.x - nreset) * 4 + (threadIdx.x >> 5));
        if (b < st.nchunks) scatter_chunk(st, b, (int)(seq & 1u) ^ 1, lane);     // the buffer the kernel before this one filled
        if (b == 0 && lane == 0) {   // hand over to the next step: nobody in this launch reads these words
            st.counter[CTR_SEQ] = seq + 1u;
            st.counter[CTR_QUEUE] = 0u;                              // work-queue head of the next step kernel
            st.counter[CTR_DONE + ((seq + 1u) & 1u)] = 0u;           // the next step's done list is empty
            if (count_step) st.counter[CTR_STEPS] += 1u;             // steps since the statistics were reset
        }
    }
}

#ifndef TVC_MIN_BLOCKS_V2
#define TVC_MIN_BLOCKS_V2 4
#endif
#ifndef TVC_V2_BLOCK
#define TVC_V2_BLOCK 128
#endif

// The hot kernel: a persistent grid (TVC_MIN_BLOCKS_V2 CTAs of TVC_V2_BLOCK threads per SM); every warp pulls 32-env groups
// of the class-ordered sequence from an atomic queue on its own and never waits for another warp.  One thread = one env for
// the whole step (K substeps, contact solve inline), state in registers between one load and one store.
//
// Same-step autoreset is deferred to close_kernel.  ~1 env in 38 ends its episode per step: resetting it where it ends made
// 1-2 lanes per warp walk through three Philox blocks, the per-episode draws and a second observation (10 % of the kernel's
// warp-instructions at 1.7 live lanes).  Instead the warp stores the terminal state and observation and appends the env to
// st.done_list (one atomic per group that has such lanes, issued under the group's stores).
// An in-place reset for small batches existed as a second template instantiation; nvcc contracted the solver's FMAs
// differently in the two, so the two launch plans differed in the last bit of contact steps -- one instantiation per
// configuration cannot.  Measured and rejected: resets and the next sequence in the kernel's own tail by warps that ran out
// of work (one launch per step, correct, but 0.21 ms: the last groups in flight took 140 us instead of 14 while idle warps
// of their CTAs polled the list -- with a poll interval of 128 ns or of 4 us alike, fewer pollers less); a per-substep CTA
// barrier that keeps a CTA's warps on the same instruction lines (won 7 % when the kernel was 92 KB of code, loses 2 % now);
// an airborne-part kernel at twice the occupancy followed by a near-ground kernel (0.150 ms against 0.125 ms at the time);
// 5 or 6 CTAs per SM with spills (0.098 / 0.102 ms against 0.089).
template <bool X, int DIV, bool FOLLOW>
__global__ void __launch_bounds__(TVC_V2_BLOCK, TVC_MIN_BLOCKS_V2)
step_kernel_v2(const __grid_constant__ DevCfg c, const __grid_constant__ DevState st, const __grid_constant__ DevIO io) {
    // launched with programmatic stream serialization: everything above this line may overlap close_kernel's tail
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int ngroups = (int)((st.n + 31) / 32);
    unsigned *const queue = &st.counter[CTR_QUEUE];
    const unsigned seq = st.counter[CTR_SEQ];    // sequence number (device-side: see close_kernel); its parity picks the buffers
    const int par = (int)(seq & 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) st.counter[CTR_SEQ2] = seq;
    const bool any_info = io.altitude || io.tilt_deg || io.omega_mag || io.fuel || io.position || io.phase || io.step || io.success ||
                          io.criteria_met || io.comp;
#ifdef TVC_PHASE_PROF2
    long long pt_prev = clock64();
#endif
    bool first_pull = true;
    int g_next = 0;
    for (;;) {
        int g = 0;
        // every warp's first group is its own index (no atomic: 2,368 warps hitting one counter at t = 0 showed up as 6 % of
        // the stall samples); after that a dynamic queue over the remaining 32-env groups of the sequence
        if (first_pull) {
            g = (int)(blockIdx.x * (TVC_V2_BLOCK / 32) + (threadIdx.x >> 5));
            first_pull = false;
        } else g = (int)(gridDim.x * (TVC_V2_BLOCK / 32)) + __shfl_sync(full, g_next, 0);   // pulled while the previous group was in its second half (below)
        if (g >= ngroups) break;
        const long long slot = (long long)g * 32 + lane;
        const bool live = slot < st.n;
        int done = 0, viol = 0;
        int ev_len = 0, ev_succ = 0, ev_reason = 0, ev_trunc = 0;
        float ev_ret = 0.0f, ev_alt = 0.0f, ev_tilt = 0.0f, ev_fuel = 0.0f;
        Env e;
        BodyP P;
        Forces f;
        long long i = 0, gid = 0;
#ifdef TVC_PHASE_PROF2
        const long long pt0 = clock64();
        Ph2 ph2s = {0u, 0u, 0u}; Ph2 *ph2 = &ph2s;
#endif
        if (live) {
            i = st.order[slot];   // the sequence is a flat array: one coalesced load
            gid = c.env_base + i;
            load_env<false>(st, X, i, e);
            float2 a;
            if (io.actions) a = io.actions[i];
            else {
                uint4 rr = philox(c.seed_lo, c.seed_hi, gid, ST_ACTION, (unsigned)io.t, (unsigned)(io.t >> 32));
                a = make_float2(2.0f * u01(rr.x) - 1.0f, 2.0f * u01(rr.y) - 1.0f);
            }
            if (io.actions_out) io.actions_out[i] = a;
            env_pre<X>(c, st, i, e, a.x, a.y, P, f);
        }
        PH2_CLK(pt1);
#ifdef TVC_SOLVE_STATS
        unsigned ss_mask = 0u;
#endif
        if (live) integrate_thread<false, FOLLOW>(c, P, e, f PH2_PASS SS_PASS);
        PH2_CLK(pt2);
#ifdef TVC_SOLVE_STATS
        {
            const long long p0 = (long long)g * 32;
            const int scls = p0 < (long long)st.totals[0] ? 0 : (p0 < (long long)st.totals[0] + st.totals[1] ? 1 : 2);
            for (int k = 0; k < c.K && k < 16; k++) {
                const unsigned m = __ballot_sync(full, live && ((ss_mask >> k) & 1u));
                if (lane == 0) {
                    atomicAdd(&g_ss[scls][k][0], 1ull);
                    if (m) { atomicAdd(&g_ss[scls][k][1], 1ull); atomicAdd(&g_ss[scls][k][2], (unsigned long long)__popc(m)); }
                }
            }
        }
#endif
        // the next group of the sequence: the atomic's round trip (~1 us with 2,368 warps on one counter) runs under this
        // group's second half instead of stalling the warp at the top of the loop
        // (ptxas turns an atomic on a warp-uniform address into its warp-aggregated form -- vote, one ATOMG, SHFL of the returned
        //  value -- and that SHFL waits for the round trip on the spot: 2 % of the kernel's stall samples sat there.  The address
        //  below is not provably uniform (threadIdx.y is 0 in this 1-D launch), so the value is first touched at the loop top.)
        if (lane == 0) {
            unsigned old;
            asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(queue + threadIdx.y));
            g_next = (int)old;      // (untouched until the loop top: any arithmetic here would wait for the round trip)
        }
        if (live) {
            StepResult r;
            env_post<X, DIV>(c, st, i, gid, e, f.a0, f.a1, r);
            viol = r.viol;
            done = r.terminated | r.truncated;
            io.reward[i] = r.reward;
            io.term[i] = (uint8_t)r.terminated;
            io.trunc[i] = (uint8_t)r.truncated;
            if (any_info) {     // tvc_step_ex's optional per-env info planes: one uniform test on the plain step path
                if (io.altitude) io.altitude[i] = r.alt;
                if (io.tilt_deg) io.tilt_deg[i] = r.tilt * 57.29577951308232f;
                if (io.omega_mag) io.omega_mag[i] = r.wmag;
                if (io.fuel) io.fuel[i] = r.fuel;
                if (io.position) { io.position[3 * i] = e.px; io.position[3 * i + 1] = e.py; io.position[3 * i + 2] = e.pz; }
                if (io.phase) io.phase[i] = e.phase;
                if (io.step) io.step[i] = e.step;
                if (io.success) io.success[i] = (uint8_t)e.success;
                if (io.criteria_met) io.criteria_met[i] = (uint8_t)(e.consec >= 10);
                if (io.comp) {
#pragma unroll
                    for (int k = 0; k < 12; k++) io.comp[12 * i + k] = r.comp[k];
                }
            }
            if (done) {
                ev_len = e.step; ev_succ = e.success; ev_reason = r.reason; ev_trunc = r.truncated;
                ev_ret = e.ep_ret; ev_alt = r.alt; ev_tilt = r.tilt; ev_fuel = r.fuel;
                if (io.final_obs) {
                    float2 *f2 = reinterpret_cast<float2 *>(io.final_obs + 10 * i);
#pragma unroll
                    for (int k = 0; k < 5; k++) f2[k] = make_float2(r.obs[2 * k], r.obs[2 * k + 1]);
                }
            }
            store_env(st, X, i, e);
            float2 *o2 = reinterpret_cast<float2 *>(io.obs + 10 * i);
#pragma unroll
            for (int k = 0; k < 5; k++) o2[k] = make_float2(r.obs[2 * k], r.obs[2 * k + 1]);
        }
        // the env's place in the NEXT sequence: class byte + (class, chunk) counters; an env whose episode ended goes to the
        // done list (close_kernel re-initialises it) and is airborne in the next sequence (a fresh env starts at z = 1 m)
        const bool relist = live && done && c.autoreset;
        const unsigned rm = __ballot_sync(full, relist);
        int rbase = 0;
        {
            if (rm && lane == 0) {      // (plain PTX for the same reason: the base is first needed after the class counters below)
                unsigned old;
                asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(&st.counter[CTR_DONE + par] + threadIdx.y), "r"((unsigned)__popc(rm)));
                rbase = (int)old;
            }
            int cl = 2;
            if (live && !relist) cl = class_of(c, X, e.pz, e.qx, e.qy, e.qz, e.qw, e.vz, e.wx, e.wy, e.wz, e.cg_off);
            if (live) st.cls[i] = (uint8_t)cl;
            count_class(st, par ^ 1, live, i, cl);   // the buffer the close_kernel of THIS step reads
        }
#ifdef TVC_PHASE_PROF2
        {
            const long long pt3 = clock64();
            const int cls = ((long long)g * 32 < (long long)st.totals[0] + st.totals[1]) ? 0 : 1;   // by position
            const unsigned tot = (unsigned)(pt2 - pt1);
            const unsigned v1 = __reduce_max_sync(full, (unsigned)(pt1 - pt0)), v3 = __reduce_max_sync(full, ph2s.setup);
            const unsigned v4 = __reduce_max_sync(full, ph2s.sweeps), v5 = __reduce_max_sync(full, (unsigned)(pt3 - pt2));
            const unsigned v7 = __reduce_max_sync(full, ph2s.bar), vt = __reduce_max_sync(full, tot);
            if (lane == 0 && g < ngroups) {
                atomicAdd(&g_ph2[cls][0], (unsigned long long)(pt0 - pt_prev));
                atomicAdd(&g_ph2[cls][1], (unsigned long long)v1);
                atomicAdd(&g_ph2[cls][2], (unsigned long long)(vt - v3 - v4 - v7));
                atomicAdd(&g_ph2[cls][3], (unsigned long long)v3); atomicAdd(&g_ph2[cls][4], (unsigned long long)v4);
                atomicAdd(&g_ph2[cls][5], (unsigned long long)v5); atomicAdd(&g_ph2[cls][6], 1ull);
                atomicAdd(&g_ph2[cls][7], (unsigned long long)v7);
            }
            pt_prev = clock64();
        }
#endif
        // episode statistics: this group owns row g of `partial` for the whole launch (deterministic).  Counts come from
        // ballots; the sums walk the (one or two) lanes whose episode ended, in lane order; lane L < 14 then holds statistic L.
        const unsigned dm = __ballot_sync(full, done), vm = __ballot_sync(full, viol);
        if (dm | vm) {
            const int n_succ = __popc(__ballot_sync(full, done && ev_succ)), n_tr = __popc(__ballot_sync(full, ev_trunc));
            const unsigned r2 = __ballot_sync(full, ev_reason == 2), r3 = __ballot_sync(full, ev_reason == 3);
            const unsigned r4 = __ballot_sync(full, ev_reason == 4), r5 = __ballot_sync(full, ev_reason == 5);
            double d_ret = 0.0, d_ret2 = 0.0, d_alt = 0.0, d_tilt = 0.0, d_fuel = 0.0;
            int n_len = 0;
            for (unsigned m = dm; m; m &= m - 1u) {
                const int src = __ffs(m) - 1;
                const double rt = (double)__shfl_sync(full, ev_ret, src);
                d_ret += rt; d_ret2 += rt * rt;
                d_alt += (double)__shfl_sync(full, ev_alt, src); d_tilt += (double)__shfl_sync(full, ev_tilt, src);
                d_fuel += (double)__shfl_sync(full, ev_fuel, src);
                n_len += __shfl_sync(full, ev_len, src);
            }
            // The row has exactly one writer per launch (this group) and every slot gets at most one add, so the order-free
            // reductions are still deterministic; as reductions they do not make the warp wait for the old value.  Every lane
            // holds every statistic (ballots and fixed-order shuffle sums), so lane 0 issues the adds one after the other behind
            // warp-uniform tests: ~10 instructions for the usual episode end (an earlier form let lane L < 14 pick statistic L with
            // a select chain over 64-bit values -- 140 instructions and 8 % of the kernel's stall samples; a switch before that
            // compiled to a 14-way divergent jump table).
            double *row = &st.partial[(long long)g * TVC_NSTAT];
            const int n_ep = __popc(dm), n_r2 = __popc(r2), n_r3 = __popc(r3), n_r4 = __popc(r4), n_r5 = __popc(r5), n_v = __popc(vm);
            if (lane == 0) {
                if (n_ep) {
                    atomicAdd(row + 0, (double)n_ep); atomicAdd(row + 1, d_ret); atomicAdd(row + 2, d_ret2); atomicAdd(row + 3, (double)n_len);
                    atomicAdd(row + 11, d_alt); atomicAdd(row + 12, d_tilt); atomicAdd(row + 13, d_fuel);
                }
                if (n_succ) atomicAdd(row + 4, (double)n_succ);
                if (n_r2) atomicAdd(row + 5, (double)n_r2);
                if (n_r3) atomicAdd(row + 6, (double)n_r3);
                if (n_r4) atomicAdd(row + 7, (double)n_r4);
                if (n_r5) atomicAdd(row + 8, (double)n_r5);
                if (n_tr) atomicAdd(row + 9, (double)n_tr);
                if (n_v) atomicAdd(row + 10, (double)n_v);
            }
        }
        if (rm) {   // the done-list slots: the base's round trip ran under the class counters and the statistics
            rbase = __shfl_sync(full, rbase, 0);
            if (relist) st.done_list[rbase + __popc(rm & ((1u << lane) - 1u))] = (int)i;
        }
    }
}

template <bool X>
__global__ void __launch_bounds__(TVC_BLOCK)
reset_kernel(const __grid_constant__ DevCfg c, const __grid_constant__ DevState st, const uint8_t *mask, float *obs,
             int first_time) {
    const long long i = (long long)blockIdx.x * TVC_BLOCK + threadIdx.x;
    if (i >= st.n) return;
    if (mask && !mask[i]) return;
    const long long gid = c.env_base + i;
    Env e;
    if (first_time) { memset(&e, 0, sizeof(e)); e.episode = -1; }
    else load_env(st, X, i, e);
    reset_env(c, X, gid, e, first_time != 0);
    store_env(st, X, i, e);
    if (obs) {
        float o[10];
        build_obs(c, X, gid, e, 0, o);
#pragma unroll
        for (int k = 0; k < 10; k++) obs[10 * i + k] = o[k];
    }
}

template <bool X>
__global__ void get_state_kernel(const __grid_constant__ DevState st, tvc_env_state *out, int delay) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    Env e;
    load_env(st, X, i, e);
    tvc_env_state s;
    memset(&s, 0, sizeof(s));
    s.pos[0] = e.px; s.pos[1] = e.py; s.pos[2] = e.pz;
    s.quat[0] = e.qx; s.quat[1] = e.qy; s.quat[2] = e.qz; s.quat[3] = e.qw;
    s.vel[0] = e.vx; s.vel[1] = e.vy; s.vel[2] = e.vz;
    s.omega[0] = e.wx; s.omega[1] = e.wy; s.omega[2] = e.wz;
    s.prev_action[0] = e.ap0; s.prev_action[1] = e.ap1;
    s.ep_return = e.ep_ret;
    s.step = e.step; s.burn = e.burn; s.phase = e.phase; s.success = e.success; s.has_prev = e.has_prev;
    s.consec = e.consec; s.hist_count = e.hist_count; s.episode = e.episode; s.n_clip = e.n_clip; s.n_run = e.n_run;
    for (int k = 0; k < 10; k++) s.ring10[k] = st.ring[12 * i + k];
    s.mass_scale = e.mass_scale; s.thrust_scale = e.thrust_scale; s.cg_offset = e.cg_off;
    s.wind[0] = e.wind_x; s.wind[1] = e.wind_y;
    if (X) for (int k = 0; k < delay; k++) { float2 d = st.delay[(long long)k * st.n + i]; s.delay_ring[k][0] = d.x; s.delay_ring[k][1] = d.y; }
    if (st.clipb) for (int k = 0; k < 32; k++) { s.clip_bits[k] = st.clipb[(long long)k * st.n + i]; s.run_bits[k] = st.runb[(long long)k * st.n + i]; }
    out[i] = s;
}

template <bool X>
__global__ void set_state_kernel(const __grid_constant__ DevState st, const tvc_env_state *in, int delay) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    const tvc_env_state s = in[i];
    Env e;
    e.px = s.pos[0]; e.py = s.pos[1]; e.pz = s.pos[2]; e.ep_ret = s.ep_return;
    e.qx = s.quat[0]; e.qy = s.quat[1]; e.qz = s.quat[2]; e.qw = s.quat[3];
    {   // the integrator relies on unit quaternions (it normalises at the end of every substep): a blob that is unit to
        // rounding is taken bit for bit, anything else is normalised on import
        const float d = e.qx * e.qx + e.qy * e.qy + e.qz * e.qz + e.qw * e.qw;
        if (fabsf(d - 1.0f) > 1e-6f && d > 0.0f) { const float sc = rsqrtf(d); e.qx *= sc; e.qy *= sc; e.qz *= sc; e.qw *= sc; }
    }
    e.vx = s.vel[0]; e.vy = s.vel[1]; e.vz = s.vel[2]; e.step = s.step;
    e.wx = s.omega[0]; e.wy = s.omega[1]; e.wz = s.omega[2];
    e.burn = s.burn; e.phase = s.phase; e.success = s.success; e.has_prev = s.has_prev; e.consec = s.consec;
    e.ap0 = s.prev_action[0]; e.ap1 = s.prev_action[1];
    e.hist_count = s.hist_count; e.n_clip = s.n_clip; e.n_run = s.n_run;
    e.mass_scale = s.mass_scale; e.thrust_scale = s.thrust_scale; e.cg_off = s.cg_offset;
    e.wind_x = s.wind[0]; e.wind_y = s.wind[1]; e.episode = s.episode;
    store_env(st, X, i, e);
    for (int k = 0; k < 10; k++) st.ring[12 * i + k] = s.ring10[k];
    if (X) for (int k = 0; k < delay; k++) st.delay[(long long)k * st.n + i] = make_float2(s.delay_ring[k][0], s.delay_ring[k][1]);
    if (st.clipb) for (int k = 0; k < 32; k++) { st.clipb[(long long)k * st.n + i] = s.clip_bits[k]; st.runb[(long long)k * st.n + i] = s.run_bits[k]; }
}

// [N][1000] caller layout <-> [1000][N] planes (TVC_DIV_EXACT window)
__global__ void hist_copy_kernel(float *planes, float *rows, long long n, int to_planes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < TVC_HIST; k++) {
        if (to_planes) planes[(long long)k * n + i] = rows[i * TVC_HIST + k];
        else rows[i * TVC_HIST + k] = planes[(long long)k * n + i];
    }
}

template <bool X>
__global__ void info_kernel(const __grid_constant__ DevState st, const __grid_constant__ DevIO io) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n) return;
    Env e;
    load_env(st, X, i, e);
    float ox, oy, oz, ow, pitch, yaw;
    reported_quat(e.qx, e.qy, e.qz, e.qw, ox, oy, oz, ow);
    euler_pitch_yaw(ox, oy, oz, ow, pitch, yaw);
    if (io.altitude) io.altitude[i] = e.pz;
    if (io.tilt_deg) io.tilt_deg[i] = sqrtf(pitch * pitch + yaw * yaw) * 57.29577951308232f;
    if (io.omega_mag) io.omega_mag[i] = sqrtf(e.wx * e.wx + e.wy * e.wy + e.wz * e.wz);
    if (io.fuel) io.fuel[i] = fuel_of(e.burn);
    if (io.position) { io.position[3 * i] = e.px; io.position[3 * i + 1] = e.py; io.position[3 * i + 2] = e.pz; }
    if (io.phase) io.phase[i] = e.phase;
    if (io.step) io.step[i] = e.step;
    if (io.success) io.success[i] = (uint8_t)e.success;
    if (io.criteria_met) io.criteria_met[i] = (uint8_t)(e.consec >= 10);
}

// deterministic reduction of the statistics rows: 512 threads = 32 row-lanes x 16 statistics (each row is one
// coalesced 128-byte read), row-strided partial sums in a fixed order, then a fixed-order fold over the row-lanes
__global__ void __launch_bounds__(512)
stats_reduce_kernel(double *partial, int nrows, double *out, unsigned *step_counter, double envs, int reset_after) {
    __shared__ double s[32][TVC_NSTAT];
    const int k = threadIdx.x & (TVC_NSTAT - 1), rl = threadIdx.x >> 4;
    double acc = 0.0;
    for (int r = rl; r < nrows; r += 32) {
        acc += partial[(long long)r * TVC_NSTAT + k];
        if (reset_after) partial[(long long)r * TVC_NSTAT + k] = 0.0;
    }
    s[rl][k] = acc;
    __syncthreads();
    if (threadIdx.x < TVC_NSTAT) {
        double t = 0.0;
        for (int j = 0; j < 32; j++) t += s[j][threadIdx.x];
        out[threadIdx.x] = (threadIdx.x == 14) ? envs * (double)step_counter[0] : t;
        if (threadIdx.x == 14 && reset_after) step_counter[0] = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Every entry point that touches CUDA runs on the handle's device and restores the caller's current device afterwards
// (several engines in one process, or an engine on a device other than the current one).
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DevGuard() { if (switched) cudaSetDevice(prev); }
};

static thread_local std::string g_err;
const char *tvc_set_err(const std::string &m) { g_err = m; return g_err.c_str(); }

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            tvc_set_err(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
            return TVC_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

// Launch with the programmatic-stream-serialization attribute: the kernel may be scheduled while its predecessor in the
// stream drains; it calls griddepcontrol.wait before touching anything the predecessor wrote.
template <typename... KArgs, typename... Args>
static cudaError_t launch_dep(void (*kernel)(KArgs...), int grid, int block, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

extern "C" {

int tvc_abi_version(void) { return TVC_ABI_VERSION; }
const char *tvc_last_error(void) { return g_err.c_str(); }

int tvc_config_default(tvc_config *c, int contract) {
    if (!c || (contract != TVC_CONTRACT_R && contract != TVC_CONTRACT_X)) { tvc_set_err("tvc_config_default: bad argument"); return TVC_E_BADARG; }
    memset(c, 0, sizeof(*c));
    c->abi_version = TVC_ABI_VERSION;
    c->contract = contract;
    c->substeps = contract == TVC_CONTRACT_R ? 4 : 10;
    c->max_episode_steps = 1000;
    c->autoreset = 0;
    c->quirks = contract == TVC_CONTRACT_R ? TVC_Q_ALL_REFERENCE : TVC_Q_CONTRACT_X;
    c->diversity_mode = contract == TVC_CONTRACT_R ? TVC_DIV_EXACT : TVC_DIV_FAST;
    c->contact_iters = 2;
    c->contact_warm_iters = 1;
    c->ground = 1;
    c->dt_step = 0.02;
    c->gradient_penalty = 0.1f; c->diversity_bonus = 0.05f;
    c->mass = 2.0f; c->radius = 0.05f; c->length = 1.0f; c->thrust = 35.0f;
    c->gimbal_max_rad = (float)(18.0 * (3.14159265358979323846 / 180.0));
    c->lin_damp = 0.01f; c->ang_damp = 0.02f;
    c->thrust_lo = 0.4f; c->thrust_hi = 1.6f;
    // contact material: enhanced_rocket_tvc_env.py:349-352 (plane) x :455-458 (rocket); Bullet combination rules
    c->contact_mu = 0.3f * 0.8f; c->contact_mu_spin = 0.1f * 0.8f + 0.1f * 0.3f; c->contact_mu_roll = 0.05f * 0.8f + 0.05f * 0.3f;
    c->contact_restitution = 0.1f; c->contact_rest_threshold = 0.2f; c->contact_erp = 0.2f; c->contact_margin = 0.05f;
    if (contract == TVC_CONTRACT_X) {
        c->mass_variation = 0.3f; c->thrust_std = 0.2f; c->cg_offset_max = 0.1f; c->wind_std = 3.0f;
        c->sensor_noise_std = 0.02f;
    }
    c->seed = 42;
    return TVC_OK;
}

}  // extern "C"

static int validate(const tvc_config *c, int64_t n) {
    if (!c) { tvc_set_err("config is NULL"); return TVC_E_BADARG; }
    if (c->abi_version != TVC_ABI_VERSION) { tvc_set_err("tvc_config.abi_version mismatch"); return TVC_E_ABI; }
    if (n <= 0 || n > (long long)INT32_MAX - TVC_CHUNK) { tvc_set_err("num_envs out of range [1, 2^31 - 1 - 1024] (env ids and the work sequence are int32)"); return TVC_E_BADARG; }
    if (c->contract != TVC_CONTRACT_R && c->contract != TVC_CONTRACT_X) { tvc_set_err("bad contract"); return TVC_E_BADARG; }
    if (c->substeps < 1 || c->substeps > 64) { tvc_set_err("substeps out of range [1,64]"); return TVC_E_BADARG; }
    if (c->max_episode_steps < 1) { tvc_set_err("max_episode_steps < 1"); return TVC_E_BADARG; }
    if (c->diversity_mode < 0 || c->diversity_mode > 2) { tvc_set_err("bad diversity_mode"); return TVC_E_BADARG; }
    if (c->delay_steps < 0 || c->delay_steps > TVC_MAX_DELAY) { tvc_set_err("delay_steps out of range"); return TVC_E_BADARG; }
    if (c->contact_iters < 0 || c->contact_iters > 256 || c->contact_warm_iters < 0 || c->contact_warm_iters > 256) { tvc_set_err("contact_iters out of range"); return TVC_E_BADARG; }
    if (!(c->dt_step > 0) || !(c->mass > 0) || !(c->radius > 0) || !(c->length > 0)) { tvc_set_err("non-positive physical parameter"); return TVC_E_BADARG; }
    // the thrust-angle sine / cosine are polynomials valid to 0.8 rad (sincos_small)
    if (!(c->gimbal_max_rad > 0.0f && c->gimbal_max_rad <= 0.8f)) { tvc_set_err("gimbal_max_rad out of range (0, 0.8]"); return TVC_E_BADARG; }
    if (!(c->mass_variation >= 0.0f && c->mass_variation < 1.0f)) { tvc_set_err("mass_variation out of range [0, 1)"); return TVC_E_BADARG; }
    if (!(c->thrust_lo <= c->thrust_hi) || !(c->thrust_std >= 0.0f)) { tvc_set_err("thrust_lo > thrust_hi or thrust_std < 0"); return TVC_E_BADARG; }
    if (!(c->propellant_fraction >= 0.0f && c->propellant_fraction < 1.0f)) { tvc_set_err("propellant_fraction out of range [0, 1)"); return TVC_E_BADARG; }
    if (!(c->contact_mu >= 0.0f) || !(c->contact_mu_spin >= 0.0f) || !(c->contact_mu_roll >= 0.0f) || !(c->contact_restitution >= 0.0f && c->contact_restitution <= 1.0f) ||
        !(c->contact_rest_threshold >= 0.0f) || !(c->contact_erp >= 0.0f && c->contact_erp <= 1.0f) || !(c->contact_margin > 0.0f)) {
        tvc_set_err("contact material parameter out of range"); return TVC_E_BADARG;
    }
    if (c->quirks & ~TVC_Q_ALL_REFERENCE) { tvc_set_err("unknown quirk bit"); return TVC_E_BADARG; }
    return TVC_OK;
}

template <typename T>
static int dalloc(T **p, size_t count) {
    CUDA_OK(cudaMalloc((void **)p, count * sizeof(T)));
    CUDA_OK(cudaMemset(*p, 0, count * sizeof(T)));
    return TVC_OK;
}

extern "C" {

int tvc_create(const tvc_config *cfg, int device, int64_t num_envs, tvc_handle **out) {
    if (!out) { tvc_set_err("out is NULL"); return TVC_E_BADARG; }
    *out = nullptr;
    int rc = validate(cfg, num_envs);
    if (rc) return rc;
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { tvc_set_err("device index out of range"); return TVC_E_BADARG; }
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        tvc_set_err(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    "; libtvc_b200 is built for sm_100a only (no fallback path)");
        return TVC_E_DEVICE;
    }
    DevGuard _guard(device);
    tvc_handle *h = new (std::nothrow) tvc_handle();
    if (!h) { tvc_set_err("out of host memory"); return TVC_E_NOMEM; }
    h->device = device; h->n = num_envs; h->base = *cfg; h->cur = *cfg; h->num_sms = prop.multiProcessorCount;
    make_devcfg(h->cur, h->dc);
    const size_t n = (size_t)num_envs;
    h->grid = (int)((num_envs + TVC_BLOCK - 1) / TVC_BLOCK);
    DevState &s = h->st;
    memset(&s, 0, sizeof(s));
    s.n = num_envs;
#define TRY(x) do { rc = (x); if (rc) { tvc_destroy(h); return rc; } } while (0)
    TRY(dalloc(&s.s0, n)); TRY(dalloc(&s.s1, n)); TRY(dalloc(&s.s2, n)); TRY(dalloc(&s.s3, n)); TRY(dalloc(&s.s4, n));
    TRY(dalloc(&s.ring, 12 * n));
    if (cfg->contract == TVC_CONTRACT_X) {
        TRY(dalloc(&s.d0, n)); TRY(dalloc(&s.d1, n));
        TRY(dalloc(&s.delay, (size_t)TVC_MAX_DELAY * n));
    }
    if (cfg->diversity_mode == TVC_DIV_FAST) { TRY(dalloc(&s.clipb, 32 * n)); TRY(dalloc(&s.runb, 32 * n)); }
    if (cfg->diversity_mode == TVC_DIV_EXACT) TRY(dalloc(&s.hist, (size_t)TVC_HIST * n));
    h->ngroups = (int)((num_envs + 31) / 32);
    TRY(dalloc(&s.partial, (size_t)h->ngroups * TVC_NSTAT));
    TRY(dalloc(&s.order, n));
    s.nchunks = (int)((num_envs + TVC_CHUNK - 1) / TVC_CHUNK);
    s.nsuper = (s.nchunks + TVC_SUPER - 1) / TVC_SUPER;
    TRY(dalloc(&s.ccount, (size_t)2 * 3 * s.nchunks));
    TRY(dalloc(&s.scount, (size_t)2 * 3 * s.nsuper));
    TRY(dalloc(&s.done_list, n));
    { cudaError_t e_ = cudaMemset(s.done_list, 0xFF, n * sizeof(int)); if (e_ != cudaSuccess) { tvc_set_err(cudaGetErrorString(e_)); tvc_destroy(h); return TVC_E_CUDA; } }   // -1 = empty slot
    TRY(dalloc(&s.totals, (size_t)4));
    TRY(dalloc(&s.counter, (size_t)CTR_WORDS));
    TRY(dalloc(&s.cls, n));
    TRY(dalloc(&h->stats_dev, (size_t)TVC_NSTAT));
    {
        cudaError_t e = cudaMallocHost((void **)&h->stats_host, sizeof(double) * TVC_NSTAT);
        if (e != cudaSuccess) { tvc_set_err(cudaGetErrorString(e)); tvc_destroy(h); return TVC_E_CUDA; }
        e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { tvc_set_err(cudaGetErrorString(e)); tvc_destroy(h); return TVC_E_CUDA; }
    }
#undef TRY
    // first-time initialisation == reset with cleared histories
    if (cfg->contract == TVC_CONTRACT_X) reset_kernel<true><<<h->grid, TVC_BLOCK>>>(h->dc, h->st, nullptr, nullptr, 1);
    else reset_kernel<false><<<h->grid, TVC_BLOCK>>>(h->dc, h->st, nullptr, nullptr, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { tvc_set_err(std::string("init kernel: ") + cudaGetErrorString(e)); tvc_destroy(h); return TVC_E_CUDA; }
    *out = h;
    return TVC_OK;
}

int tvc_destroy(tvc_handle *h) {
    if (!h) return TVC_OK;
    DevGuard _guard(h->device);
    DevState &s = h->st;
    cudaFree(s.s0); cudaFree(s.s1); cudaFree(s.s2); cudaFree(s.s3); cudaFree(s.s4);
    cudaFree(s.d0); cudaFree(s.d1); cudaFree(s.ring); cudaFree(s.clipb); cudaFree(s.runb); cudaFree(s.hist);
    cudaFree(s.delay); cudaFree(s.partial); cudaFree(s.order); cudaFree(s.ccount); cudaFree(s.scount); cudaFree(s.done_list); cudaFree(s.totals); cudaFree(s.counter); cudaFree(s.cls); cudaFree(h->stats_dev);
    cudaFree(h->io_act); cudaFree(h->io_obs) /* the obs|reward|flags slab */; cudaFree(h->io_final);
    tvc_rollout_free(h);
    tvc_curiosity_free(h);
    if (h->stats_host) cudaFreeHost(h->stats_host);
    if (h->act_pinned) cudaFreeHost(h->act_pinned);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return TVC_OK;
}

#define CHECK_H(h) if (!(h)) { tvc_set_err("handle is NULL"); return TVC_E_BADARG; } DevGuard _guard((h)->device)
#define LAUNCH_OK(what) do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { tvc_set_err(std::string(what) + ": " + cudaGetErrorString(_e)); return TVC_E_CUDA; } } while (0)

int tvc_reset(tvc_handle *h, const uint8_t *mask_dev, uint64_t seed, float *obs_out_dev, tvc_stream stream) {
    CHECK_H(h);
    if (seed != TVC_SEED_KEEP) { h->cur.seed = seed; h->base.seed = seed; h->dc.seed_lo = (unsigned)seed; h->dc.seed_hi = (unsigned)(seed >> 32); }
    cudaStream_t s = (cudaStream_t)stream;
    if (h->cur.contract == TVC_CONTRACT_X) reset_kernel<true><<<h->grid, TVC_BLOCK, 0, s>>>(h->dc, h->st, mask_dev, obs_out_dev, 0);
    else reset_kernel<false><<<h->grid, TVC_BLOCK, 0, s>>>(h->dc, h->st, mask_dev, obs_out_dev, 0);
    LAUNCH_OK("reset_kernel");
    h->order_valid = false;
    return TVC_OK;
}

static int launch_step(tvc_handle *h, const DevIO &io, cudaStream_t s) {
#ifdef TVC_DBG
    { const char *e = getenv("TVC_DBG_ONLY"); h->dc.dbg_only = e ? atoi(e) : -1; }
#endif
    const bool X = h->cur.contract == TVC_CONTRACT_X;
    const int dv = h->cur.diversity_mode;
    {
        const int cgrid = h->st.nchunks;
        if (!h->order_valid) {   // first step, or the state was changed behind the step path's back: classes from the state planes
            CUDA_OK(cudaMemsetAsync(h->st.ccount, 0, sizeof(int) * 2 * 3 * (size_t)h->st.nchunks, s));
            CUDA_OK(cudaMemsetAsync(h->st.scount, 0, sizeof(int) * 2 * 3 * (size_t)h->st.nsuper, s));
            const int g256 = (int)((h->n + 255) / 256);
            if (X) classify_state_kernel<true><<<g256, 256, 0, s>>>(h->dc, h->st);
            else classify_state_kernel<false><<<g256, 256, 0, s>>>(h->dc, h->st);
            LAUNCH_OK("classify_state_kernel");
            if (X) close_kernel<true><<<(cgrid + 3) / 4, 128, 0, s>>>(h->dc, h->st, nullptr, 0, 0);
            else close_kernel<false><<<(cgrid + 3) / 4, 128, 0, s>>>(h->dc, h->st, nullptr, 0, 0);
            LAUNCH_OK("close_kernel (from state)");
        }
        if (h->v2_grid == 0) {   // persistent grid: resident CTAs of the v2 kernel, capped by the number of groups
            int per_sm = 0;
            cudaError_t e = X ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel_v2<true, 1, false>, TVC_V2_BLOCK, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel_v2<false, 1, false>, TVC_V2_BLOCK, 0);
            if (e != cudaSuccess || per_sm < 1) per_sm = 1;
            { const char *p = getenv("TVC_V2_CTAS_PER_SM"); if (p && atoi(p) >= 1 && atoi(p) < per_sm) per_sm = atoi(p); }   // diagnostics: fewer resident warps
            const int cap = per_sm * h->num_sms;
            const int wpb = TVC_V2_BLOCK / 32;
            const int ctas = (h->ngroups + wpb - 1) / wpb;
            { const char *p = getenv("TVC_PDL"); h->pdl = !(p && p[0] == '0'); }   // diagnostics: TVC_PDL=0 launches plainly
            const int want = h->ngroups <= cap ? h->ngroups : ctas;
            h->v2_grid = want < cap ? want : cap;
        }
        const bool follow = !(h->cur.quirks & TVC_Q_FROZEN_FORCES);   // quirk Q3 cleared: the thrust follows the body every substep
#define GO(XX, DD) do { if (follow) (void)launch_dep(step_kernel_v2<XX, DD, true>, h->v2_grid, TVC_V2_BLOCK, s, h->pdl, h->dc, h->st, io); \
                        else (void)launch_dep(step_kernel_v2<XX, DD, false>, h->v2_grid, TVC_V2_BLOCK, s, h->pdl, h->dc, h->st, io); } while (0)
        if (X) { if (dv == 0) GO(true, 0); else if (dv == 1) GO(true, 1); else GO(true, 2); }
        else   { if (dv == 0) GO(false, 0); else if (dv == 1) GO(false, 1); else GO(false, 2); }
        {   // close the step: reset CTAs (one env of the done list per thread for twice the expected n / 38 finished episodes; more
            // are walked grid-stride) + one warp per chunk of the next sequence, one launch
            int nreset = 0;
            if (h->cur.autoreset) { nreset = (int)((h->n + 16 * 128 - 1) / (16 * 128)); if (nreset > 4 * h->num_sms) nreset = 4 * h->num_sms; }
            if (X) (void)launch_dep(close_kernel<true>, nreset + (cgrid + 3) / 4, 128, s, h->pdl, h->dc, h->st, io.obs, nreset, 1);
            else (void)launch_dep(close_kernel<false>, nreset + (cgrid + 3) / 4, 128, s, h->pdl, h->dc, h->st, io.obs, nreset, 1);
            LAUNCH_OK("close_kernel (end of step)");
        }
        h->order_valid = true;
#undef GO
    }
    h->lifetime_steps += 1;
    return TVC_OK;
}

int tvc_step_ex(tvc_handle *h, const tvc_step_io *u, tvc_stream stream) {
    CHECK_H(h);
    if (!u || !u->obs || !u->reward || !u->terminated || !u->truncated) { tvc_set_err("tvc_step: obs/reward/terminated/truncated must be non-NULL"); return TVC_E_BADARG; }
    DevIO io;
    memset(&io, 0, sizeof(io));
    io.actions = (const float2 *)u->actions; io.obs = u->obs; io.reward = u->reward; io.term = u->terminated; io.trunc = u->truncated;
    io.final_obs = u->final_obs; io.actions_out = (float2 *)u->actions_out;
    io.altitude = u->info.altitude; io.tilt_deg = u->info.tilt_deg; io.omega_mag = u->info.omega_mag; io.fuel = u->info.fuel;
    io.position = u->info.position; io.phase = u->info.phase; io.step = u->info.step; io.success = u->info.success;
    io.criteria_met = u->info.criteria_met; io.comp = u->info.reward_components;
    io.t = (unsigned long long)h->lifetime_steps;
    return launch_step(h, io, (cudaStream_t)stream);
}

int tvc_step(tvc_handle *h, const float *actions_dev, float *obs_dev, float *reward_dev, uint8_t *terminated_dev,
             uint8_t *truncated_dev, float *final_obs_dev, tvc_stream stream) {
    tvc_step_io u;
    memset(&u, 0, sizeof(u));
    u.actions = actions_dev; u.obs = obs_dev; u.reward = reward_dev; u.terminated = terminated_dev; u.truncated = truncated_dev;
    u.final_obs = final_obs_dev;
    return tvc_step_ex(h, &u, stream);
}

int tvc_host_sync(tvc_handle *h) {
    CHECK_H(h);
    CUDA_OK(cudaStreamSynchronize(h->own_stream));
    h->host_pending = false;
    return TVC_OK;
}

int tvc_step_host(tvc_handle *h, const float *actions_host, float *obs_host, float *reward_host, uint8_t *terminated_host,
                  uint8_t *truncated_host, float *final_obs_host) {
    int rc = tvc_step_host_async(h, actions_host, obs_host, reward_host, terminated_host, truncated_host, final_obs_host);
    if (rc) return rc;
    return tvc_host_sync(h);
}

int tvc_step_host_async(tvc_handle *h, const float *actions_host, float *obs_host, float *reward_host, uint8_t *terminated_host,
                        uint8_t *truncated_host, float *final_obs_host) {
    CHECK_H(h);
    if (!obs_host || !reward_host || !terminated_host || !truncated_host) { tvc_set_err("tvc_step_host: NULL output"); return TVC_E_BADARG; }
    const size_t n = (size_t)h->n;
    if (h->host_pending) CUDA_OK(cudaStreamSynchronize(h->own_stream));   // a second enqueue without tvc_host_sync: drain first
    h->host_pending = true;
    if (!h->io_obs) {   // one-time staging buffers (not on the steady-state step path)
        int rc;
        if ((rc = dalloc(&h->io_act, 2 * n))) return rc;
        // obs | reward | terminated | truncated in ONE slab, so that a caller whose host buffers are laid out the same way
        // gets a single device-to-host copy of 46 bytes per env
        uint8_t *slab = nullptr;
        if ((rc = dalloc(&slab, 46 * n))) return rc;
        h->io_obs = (float *)slab; h->io_rew = (float *)(slab + 40 * n); h->io_term = slab + 44 * n; h->io_trunc = slab + 45 * n;
    }
    if (final_obs_host && !h->io_final) {   // only needed for pageable final_obs buffers; allocated up front to keep the path simple
        int rc;
        if ((rc = dalloc(&h->io_final, 10 * n))) return rc;
    }
    cudaStream_t s = h->own_stream;
    if (actions_host) {
        // pinned caller memory goes to the copy engine as it is; pageable memory is staged through the handle's own
        // pinned buffer first (an asynchronous copy from pageable memory would stage inside the driver anyway).
        // Measured and rejected: the step kernel reading pinned actions in place over PCIe (8 B per env beside the state
        // loads, no host-to-device copy): 0.474 ms per end-to-end step against 0.370 with the copy -- 2,368 warps' 256-byte
        // reads arrive far below the copy engine's 56 GB/s and sit at the head of every env's dependent chain.
        const float *src = actions_host;
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, actions_host) == cudaSuccess && at.type == cudaMemoryTypeHost;
        if (!pinned) {
            (void)cudaGetLastError();
            if (!h->act_pinned) CUDA_OK(cudaMallocHost((void **)&h->act_pinned, sizeof(float) * 2 * n));
            memcpy(h->act_pinned, actions_host, sizeof(float) * 2 * n);
            src = h->act_pinned;
        }
        CUDA_OK(cudaMemcpyAsync(h->io_act, src, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, s));
    }
    // Final observations exist only for the ~3 % of envs whose episode ended in this step.  When the caller's buffer is
    // pinned (device-mapped) host memory the kernel stores those rows straight into it over PCIe and the dense
    // [N,10] device-to-host copy (40 B per env) disappears; pageable memory takes the staged copy below.
    float *final_dev = nullptr;
    bool final_direct = false;
    if (final_obs_host) {
        if (final_obs_host != h->final_host_seen) {
            cudaPointerAttributes at;
            h->final_host_seen = final_obs_host;
            h->final_host_dev = nullptr;
            if (cudaPointerGetAttributes(&at, final_obs_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
                h->final_host_dev = (float *)at.devicePointer;
            else (void)cudaGetLastError();
        }
        final_direct = h->final_host_dev != nullptr;
        final_dev = final_direct ? h->final_host_dev : h->io_final;
    }
    int rc = tvc_step(h, actions_host ? h->io_act : nullptr, h->io_obs, h->io_rew, h->io_term, h->io_trunc, final_dev, (tvc_stream)s);
    if (rc) return rc;
    const uint8_t *ob = (const uint8_t *)obs_host;
    if ((const uint8_t *)reward_host == ob + 40 * n && terminated_host == ob + 44 * n && truncated_host == ob + 45 * n) {
        CUDA_OK(cudaMemcpyAsync(obs_host, h->io_obs, 46 * n, cudaMemcpyDeviceToHost, s));
    } else {
        CUDA_OK(cudaMemcpyAsync(obs_host, h->io_obs, sizeof(float) * 10 * n, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaMemcpyAsync(reward_host, h->io_rew, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaMemcpyAsync(terminated_host, h->io_term, n, cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaMemcpyAsync(truncated_host, h->io_trunc, n, cudaMemcpyDeviceToHost, s));
    }
    if (final_obs_host && !final_direct) CUDA_OK(cudaMemcpyAsync(final_obs_host, h->io_final, sizeof(float) * 10 * n, cudaMemcpyDeviceToHost, s));
    return TVC_OK;
}

size_t tvc_state_bytes(const tvc_handle *h) { return h ? (size_t)h->n * sizeof(tvc_env_state) : 0; }

int tvc_get_state(tvc_handle *h, void *dev_blob, size_t bytes, tvc_stream stream) {
    CHECK_H(h);
    if (!dev_blob || bytes < tvc_state_bytes(h)) { tvc_set_err("tvc_get_state: blob too small"); return TVC_E_BADARG; }
    const int g = (int)((h->n + 127) / 128);
    if (h->cur.contract == TVC_CONTRACT_X) get_state_kernel<true><<<g, 128, 0, (cudaStream_t)stream>>>(h->st, (tvc_env_state *)dev_blob, TVC_MAX_DELAY);
    else get_state_kernel<false><<<g, 128, 0, (cudaStream_t)stream>>>(h->st, (tvc_env_state *)dev_blob, 0);
    LAUNCH_OK("get_state_kernel");
    return TVC_OK;
}

int tvc_set_state(tvc_handle *h, const void *dev_blob, size_t bytes, tvc_stream stream) {
    CHECK_H(h);
    if (!dev_blob || bytes < tvc_state_bytes(h)) { tvc_set_err("tvc_set_state: blob too small"); return TVC_E_BADARG; }
    const int g = (int)((h->n + 127) / 128);
    if (h->cur.contract == TVC_CONTRACT_X) set_state_kernel<true><<<g, 128, 0, (cudaStream_t)stream>>>(h->st, (const tvc_env_state *)dev_blob, TVC_MAX_DELAY);
    else set_state_kernel<false><<<g, 128, 0, (cudaStream_t)stream>>>(h->st, (const tvc_env_state *)dev_blob, 0);
    LAUNCH_OK("set_state_kernel");
    h->order_valid = false;
    return TVC_OK;
}

int tvc_get_reward_history(tvc_handle *h, float *dev_out, size_t bytes, tvc_stream stream) {
    CHECK_H(h);
    if (!h->st.hist) { tvc_set_err("tvc_get_reward_history: the handle does not keep the window (diversity_mode != TVC_DIV_EXACT)"); return TVC_E_STATE; }
    if (!dev_out || bytes < sizeof(float) * TVC_HIST * (size_t)h->n) { tvc_set_err("tvc_get_reward_history: buffer too small"); return TVC_E_BADARG; }
    hist_copy_kernel<<<(int)((h->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->st.hist, dev_out, h->n, 0);
    LAUNCH_OK("hist_copy_kernel");
    return TVC_OK;
}

int tvc_set_reward_history(tvc_handle *h, const float *dev_in, size_t bytes, tvc_stream stream) {
    CHECK_H(h);
    if (!h->st.hist) { tvc_set_err("tvc_set_reward_history: the handle does not keep the window (diversity_mode != TVC_DIV_EXACT)"); return TVC_E_STATE; }
    if (!dev_in || bytes < sizeof(float) * TVC_HIST * (size_t)h->n) { tvc_set_err("tvc_set_reward_history: buffer too small"); return TVC_E_BADARG; }
    hist_copy_kernel<<<(int)((h->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->st.hist, const_cast<float *>(dev_in), h->n, 1);
    LAUNCH_OK("hist_copy_kernel");
    return TVC_OK;
}

int tvc_read_info(tvc_handle *h, const tvc_info_soa *u, tvc_stream stream) {
    CHECK_H(h);
    if (!u) { tvc_set_err("info is NULL"); return TVC_E_BADARG; }
    DevIO io;
    memset(&io, 0, sizeof(io));
    io.altitude = u->altitude; io.tilt_deg = u->tilt_deg; io.omega_mag = u->omega_mag; io.fuel = u->fuel; io.position = u->position;
    io.phase = u->phase; io.step = u->step; io.success = u->success; io.criteria_met = u->criteria_met;
    const int g = (int)((h->n + 127) / 128);
    if (h->cur.contract == TVC_CONTRACT_X) info_kernel<true><<<g, 128, 0, (cudaStream_t)stream>>>(h->st, io);
    else info_kernel<false><<<g, 128, 0, (cudaStream_t)stream>>>(h->st, io);
    LAUNCH_OK("info_kernel");
    return TVC_OK;
}

int tvc_episode_stats_dev(tvc_handle *h, double *dev_out, int reset_after, tvc_stream stream) {
    CHECK_H(h);
    if (!dev_out) { tvc_set_err("dev_out is NULL"); return TVC_E_BADARG; }
    // the step count lives on the device (bumped by the kernel that closes a step), so that CUDA-graph replays count too
    stats_reduce_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(h->st.partial, h->ngroups, dev_out, h->st.counter + CTR_STEPS, (double)h->n, reset_after);
    LAUNCH_OK("stats_reduce_kernel");
    return TVC_OK;
}

int tvc_episode_stats(tvc_handle *h, double *host_out, int reset_after, tvc_stream stream) {
    CHECK_H(h);
    if (!host_out) { tvc_set_err("host_out is NULL"); return TVC_E_BADARG; }
    int rc = tvc_episode_stats_dev(h, h->stats_dev, reset_after, stream);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(h->stats_host, h->stats_dev, sizeof(double) * TVC_NSTAT, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(host_out, h->stats_host, sizeof(double) * TVC_NSTAT);
    return TVC_OK;
}

int tvc_set_curriculum(tvc_handle *h, const tvc_stage_conditions *c) {
    CHECK_H(h);
    if (!c) { tvc_set_err("conditions is NULL"); return TVC_E_BADARG; }
    if (h->cur.contract != TVC_CONTRACT_X) { tvc_set_err("tvc_set_curriculum: the reference env has no curriculum coupling; use Contract X"); return TVC_E_STATE; }
    tvc_config n = h->base;
    n.init_tilt_max = c->max_initial_tilt;
    n.init_omega_max = c->max_initial_angular_vel;
    if (!c->domain_randomization) { n.mass_variation = 0.0f; n.thrust_std = 0.0f; n.cg_offset_max = 0.0f; }
    else n.mass_variation = c->mass_variation;
    if (!c->sensor_noise) n.sensor_noise_std = 0.0f;
    n.wind_std = c->wind_enabled ? c->wind_force : 0.0f;
    if (c->max_gimbal_angle_deg > 0.0f) n.gimbal_max_rad = c->max_gimbal_angle_deg * (float)(3.14159265358979323846 / 180.0);
    n.seed = h->cur.seed;
    {   // the same checks as at creation, before the new conditions are committed
        const int rc = validate(&n, h->n);
        if (rc) return rc;
    }
    h->cur = n;
    make_devcfg(h->cur, h->dc);
    h->order_valid = false;
    return TVC_OK;
}

int tvc_get_config(const tvc_handle *h, tvc_config *out) {
    CHECK_H(h);
    if (!out) { tvc_set_err("out is NULL"); return TVC_E_BADARG; }
    *out = h->cur;
    return TVC_OK;
}
#ifdef TVC_SOLVE_STATS
int tvc_debug_solve_stats(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tvc::g_ss, sizeof(unsigned long long) * 3 * 16 * 3);
    if (reset) { unsigned long long z[3 * 16 * 3] = {0}; cudaMemcpyToSymbol(tvc::g_ss, z, sizeof(z)); }
    return 0;
}
#endif
#ifdef TVC_PHASE_PROF2
int tvc_debug_phase2(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tvc::g_ph2, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tvc::g_ph2, z, sizeof(z)); }
    return 0;
}
#endif
#ifdef TVC_PHASE_PROF
int tvc_debug_phase(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, tvc::g_phase, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(tvc::g_phase, z, sizeof(z)); }
    return 0;
}
#endif
int64_t tvc_num_envs(const tvc_handle *h) { return h ? h->n : -1; }
int64_t tvc_lifetime_steps(const tvc_handle *h) { return h ? h->lifetime_steps : -1; }

}  // extern "C"
