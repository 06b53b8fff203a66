// transport_v9.cuh -- role-specialised photon transport (included by b200rt.cu after the shared helpers).
//
// EXPERIMENT, measured slower than the per-warp-pool kernel in every variant (profiles/ab_v9_r02.txt);
// built only with -DB200RT_WITH_V9.  Kept compiling against the current helpers (tallies always take the global path
// here); its GPU validation dates from commit ebb7d38.
//
// One block per SM.  The photon pool belongs to the BLOCK (NPB slots, structure of arrays in shared memory) and the
// warps are specialised:
//   geometry warps (the first V9_NWF warps, V9_RF registers each after `setmaxnreg.dec`): the flight phase only --
//       majorant-grid DDA with null-collision budget, no RNG, no 3-D field look-ups.  Their whole instruction stream
//       is one loop, so it stays in the instruction cache, and the small register budget buys resident warps for the
//       dependent majorant gathers that bound the phase;
//   event warps (the remaining V9_NWE warps, V9_RE registers each after `setmaxnreg.inc`): regeneration, tentative
//       collisions, scattering / surface events with their local estimates.
// Photons move between the roles through five block-level multi-producer / multi-consumer rings of slot numbers
// (DEAD, FLY, TENTATIVE, COLLISION, SURFACE).  A ring cell is 16 bits: 11 bits of slot number and a 5-bit state
// (lap number and a full bit), the bounded-queue protocol in which a writer waits for "free in my lap" and a reader
// for "full in my lap"; positions are reserved by one shared-memory atomic per WARP and ring operation.  A ring can
// never overflow: every slot is in at most one ring and the capacity is >= NPB.
// Physics, RNG streams and tallies are those of transport_kernel (v8): a photon's history depends only on its
// (job seed, global photon index), never on the warp or role that happens to process it.
#pragma once

#ifndef V9_NWF
#define V9_NWF 16
#endif
#ifndef V9_NWE
#define V9_NWE 12
#endif
#ifndef V9_RF
#define V9_RF 56
#endif
#ifndef V9_RE
#define V9_RE 88
#endif
#ifndef V9_MINB
#define V9_MINB 32          // a warp that finds fewer items than this waits (asleep) up to V9_WAIT cycles for a full batch
#endif
#ifndef V9_WAIT
#define V9_WAIT 30000       // ~16 us: a partial batch costs the issue slots of a full one, so the role with spare capacity waits
#endif
// ordering of "photon record written" before "ring cell written" (and cell read before record read), both in shared
// memory: 0 = compiler barrier only (shared-memory accesses of a warp are performed in program order by the LSU),
// 1 = fence.acq_rel.cta, 2 = fence.sc.cta (__threadfence_block)
#ifndef V9_FENCE
#define V9_FENCE 1
#endif
__device__ __forceinline__ void v9_fence() {
#if V9_FENCE == 2
    v9_fence();
#elif V9_FENCE == 1
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#else
    asm volatile("" ::: "memory");
#endif
}
#ifndef V9_NPB
#define V9_NPB 1536         // photon slots per block
#endif
#define V9_NT (32 * (V9_NWF + V9_NWE))
#define V9_CAP 2048
#define V9_SLOT_MASK 0x7ffu
static_assert(V9_NWF % 4 == 0 && V9_NWE % 4 == 0, "setmaxnreg works on warpgroups of four warps");
// registers per thread at launch: what __launch_bounds__(V9_NT, 1) lets ptxas use (it uses all of it).  `setmaxnreg.inc`
// draws ONLY on what the `setmaxnreg.dec` of the same block has released (the CTA pool) -- registers of the SM that the
// launch left unallocated do not count -- so the two sides must balance or the last event warpgroup waits forever.
#define V9_RL ((65536 / V9_NT) / 8 * 8)
#ifndef V9_NO_SETMAXNREG
static_assert(V9_RF <= V9_RL && V9_RE >= V9_RL, "geometry warps release registers, event warps acquire them");
static_assert(V9_NWF * (V9_RL - V9_RF) >= V9_NWE * (V9_RE - V9_RL), "setmaxnreg: released registers must cover the acquired ones");
#endif

enum { RQ_DEAD = 0, RQ_FLY = 1, RQ_TENT = 2, RQ_COLL = 3, RQ_SFC = 4, RQ_N = 5 };
struct V9Ctl {
    unsigned head[8];
    unsigned tail[8];
    unsigned retired;      // slots that will never be used again (photon source exhausted)
    unsigned exhausted;
    unsigned pad[2];
};
#define V9_RING_BYTES (RQ_N * V9_CAP * 2)
// 32-bit words of shared memory for pool + rings + control
#define V9_POOL_WORDS(npb) (NFIELD * (npb) + V9_RING_BYTES / 4 + int(sizeof(V9Ctl) / 4))

#ifndef V9_NO_SETMAXNREG
template <int N>
__device__ __forceinline__ void v9_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void v9_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
#else
template <int N>
__device__ __forceinline__ void v9_reg_dec() {}
template <int N>
__device__ __forceinline__ void v9_reg_inc() {}
#endif

__device__ __forceinline__ unsigned v9_ld_volatile(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

// items waiting in a ring (may be stale by the time it is used; only a hint for the phase choice)
__device__ __forceinline__ int v9_avail(const V9Ctl* c, int r) {
    const unsigned h = v9_ld_volatile(&c->head[r]);
    const unsigned t = v9_ld_volatile(&c->tail[r]);
    return int(t - h);
}

// warp-collective push: lanes with `pred` append `slot`.  The caller has fenced its pool stores.
__device__ __forceinline__ void v9_push(unsigned short* ring, unsigned* tailp, bool pred, int slot, int lane, unsigned lt_mask) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    unsigned pos = 0;
    if (lane == leader) pos = atomicAdd(tailp, unsigned(__popc(m)));
    pos = __shfl_sync(0xffffffffu, pos, leader);
    if (pred) {
        const unsigned q = pos + unsigned(__popc(m & lt_mask));
        const unsigned lap = (q / V9_CAP) & 15u;
        volatile unsigned short* c = ring + (q & (V9_CAP - 1));
        const unsigned want_free = (lap << 1) << 11;
        // the reader of the previous lap is done (never spins in practice; bounded so that a protocol error cannot hang the GPU)
        for (unsigned spin = 0; (unsigned(*c) & 0xf800u) != want_free && spin < (1u << 28); ++spin) {}
        *c = (unsigned short)((((lap << 1) | 1u) << 11) | unsigned(slot));
    }
}

// warp-collective pop of up to 32 items; returns the number taken, lane i < n gets its slot
__device__ __forceinline__ int v9_pop(unsigned short* ring, unsigned* headp, const unsigned* tailp, int lane, int& slot) {
    unsigned h = 0;
    int n = 0;
    if (lane == 0) {
        for (;;) {
            h = v9_ld_volatile(headp);
            const unsigned t = v9_ld_volatile(tailp);
            n = min(32, int(t - h));
            if (n <= 0) { n = 0; break; }
            if (atomicCAS(headp, h, h + unsigned(n)) == h) break;
        }
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    h = __shfl_sync(0xffffffffu, h, 0);
    slot = 0;
    if (lane < n) {
        const unsigned q = h + unsigned(lane);
        const unsigned lap = (q / V9_CAP) & 15u;
        volatile unsigned short* c = ring + (q & (V9_CAP - 1));
        const unsigned want_full = ((lap << 1) | 1u) << 11;
        unsigned v = unsigned(*c);                              // the writer that reserved this cell may not have stored yet
        for (unsigned spin = 0; (v & 0xf800u) != want_full && spin < (1u << 28); ++spin) v = unsigned(*c);
        slot = int(v & V9_SLOT_MASK);
        *c = (unsigned short)((((lap + 1u) & 15u) << 1) << 11);
    }
    v9_fence();                     // the slot's record is read after its ring cell
    return n;
}

// PL: flux / heating target; FZ: column-frozen photons possible; NPB: photon slots of the block; CAM: all-sky camera;
// UZ: 3-D layers equally thick and runs of empty cells enabled (slab follows from the height)
template <bool PL, bool FZ, int NPB, bool CAM, bool UZ>
__global__ void __launch_bounds__(V9_NT, 1) transport_v9(const __grid_constant__ DevScene S) {
    extern __shared__ float4 smem_f4[];
    Smem sm;
    float* pool;
    unsigned short* rings;
    V9Ctl* ctl;
    {
        float4* q4 = smem_f4;
        float4* slabA = q4; q4 += S.nslab_z;
        int4* slabB = reinterpret_cast<int4*>(q4); q4 += S.nslab_z;
        float4* grpA = q4; q4 += S.ngroup;
        int4* grpB = reinterpret_cast<int4*>(q4); q4 += S.ngroup;
        double* acc = reinterpret_cast<double*>(q4);
        double* acc_atm = acc + 4 * 32;
        double* tal = acc_atm + blockDim.x;
        const int ntal = PL ? S.ntal_flux_smem + S.ntal_heat_smem : 0;
        unsigned* cnt = reinterpret_cast<unsigned*>(tal + ntal);
        float* q = reinterpret_cast<float*>(cnt + 8 * 32);
        float* z = q; q += S.nz + 1;
        float* e1tot = q; q += S.nz;
        float* e1cum = q; q += S.nz + 1;
        float* e1 = q; q += S.np1d * S.nz;
        float* o1 = q; q += S.np1d * S.nz;
        float* a1 = q; q += S.np1d * S.nz;
        pool = q; q += NFIELD * NPB;
        rings = reinterpret_cast<unsigned short*>(q); q += V9_RING_BYTES / 4;
        ctl = reinterpret_cast<V9Ctl*>(q);
        for (int i = threadIdx.x; i <= S.nz; i += blockDim.x) { z[i] = S.zgrd[i]; e1cum[i] = S.e1cum[i]; }
        for (int i = threadIdx.x; i < S.nz; i += blockDim.x) e1tot[i] = S.e1tot[i];
        for (int i = threadIdx.x; i < S.np1d * S.nz; i += blockDim.x) { e1[i] = S.e1[i]; o1[i] = S.o1[i]; a1[i] = S.a1[i]; }
        for (int i = threadIdx.x; i < S.nslab_z; i += blockDim.x) {
            const int l0 = S.slab_lay0[i], l1 = S.slab_lay0[i + 1];
            const int cz = S.slab_cz[i];
            const int w = cz >= 0 ? (cz | (S.slab_cg[i] << 16)) : int(0x80000000u | (unsigned(S.slab_cg[i]) << 16));
            slabA[i] = make_float4(S.zgrd[l0], S.zgrd[l1], S.slab_maj1d[i], __int_as_float(w));
            slabB[i] = make_int4(l0, l1, S.slab_cg[i], 0);
        }
        for (int i = threadIdx.x; i < S.ngroup; i += blockDim.x) {
            const int s0 = S.group_lo[i], s1 = S.group_lo[i + 1];
            const int l0 = S.slab_lay0[s0], l1 = S.slab_lay0[s1];
            grpA[i] = make_float4(S.zgrd[l0], S.zgrd[l1], S.group_maj1d[i], __int_as_float(int(unsigned(s0) | (unsigned(s1) << 16))));
            grpB[i] = make_int4(s0, s1, l0, l1);
        }
        if (threadIdx.x < 32) {
            for (int k = 0; k < 4; ++k) acc[k * 32 + threadIdx.x] = 0.0;
            for (int k = 0; k < 8; ++k) cnt[k * 32 + threadIdx.x] = 0u;
        }
        acc_atm[threadIdx.x] = 0.0;
        for (int i = threadIdx.x; i < ntal; i += blockDim.x) tal[i] = 0.0;
        sm.ftal = (PL && S.ntal_flux_smem > 0) ? tal : nullptr;
        sm.htal = (PL && S.ntal_heat_smem > 0) ? tal + S.ntal_flux_smem : nullptr;
        // rings: every cell free in lap 0, except that the DEAD ring starts with all NPB slots
        for (int i = threadIdx.x; i < RQ_N * V9_CAP; i += blockDim.x) {
            const int r = i / V9_CAP, c = i - r * V9_CAP;
            rings[i] = (r == RQ_DEAD && c < NPB) ? (unsigned short)((1u << 11) | unsigned(c)) : (unsigned short)0;
        }
        if (threadIdx.x < 8) { ctl->head[threadIdx.x] = 0u; ctl->tail[threadIdx.x] = (threadIdx.x == RQ_DEAD) ? unsigned(NPB) : 0u; }
        if (threadIdx.x == 0) { ctl->retired = 0u; ctl->exhausted = 0u; }
        sm.z = z; sm.e1tot = e1tot; sm.e1cum = e1cum; sm.e1 = e1; sm.o1 = o1; sm.a1 = a1;
        sm.slabA = slabA; sm.slabB = slabB; sm.grpA = grpA; sm.grpB = grpB; sm.acc = acc; sm.acc_atm = acc_atm; sm.cnt = cnt;
    }
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool want_flux = PL && (S.target & B200RT_TARGET_FLUX) != 0;
    const bool want_rad = (S.target & B200RT_TARGET_RADIANCE) != 0 && S.nrad > 0;
    const bool want_heat = PL && (S.target & B200RT_TARGET_HEATING) != 0;
    unsigned short* const rD = rings + RQ_DEAD * V9_CAP;
    unsigned short* const rF = rings + RQ_FLY * V9_CAP;
    unsigned short* const rE = rings + RQ_TENT * V9_CAP;
    unsigned short* const rC = rings + RQ_COLL * V9_CAP;
    unsigned short* const rS = rings + RQ_SFC * V9_CAP;
#define BCAST(x) __shfl_sync(FULL, (x), 0)
#define PUSH(r, id, cond, slot) v9_push((r), &ctl->tail[id], (cond), (slot), lane, lt_mask)
#define POP(r, id, slot) v9_pop((r), &ctl->head[id], &ctl->tail[id], lane, (slot))
// idle path: sleep; a warp that has found nothing to do for ~4 s of SM clocks declares the block hung (sets the flag the
// host checks, releases every warp of the block) -- a protocol error must never cost a GPU
#define V9_IDLE()                                                                                     \
    {                                                                                                 \
        const long long now_ = __shfl_sync(FULL, clock64(), 0);                                       \
        if (idle0 == 0) idle0 = now_;                                                                 \
        else if (now_ - idle0 > (8ll << 30)) {                                                        \
            if (lane == 0) { atomicAdd(&S.stats->hang, 1ull); atomicAdd(&ctl->retired, unsigned(NPB)); } \
            break;                                                                                    \
        }                                                                                             \
        __nanosleep(200);                                                                             \
    }
#define RNG4(out)                                                                                     \
    {                                                                                                 \
        const unsigned long long seed_ = S.jobs[p.job].seed;                                          \
        out = philox_u01x4(p.rc0, p.rc1, p.rc2, unsigned(seed_), unsigned(seed_ >> 32));              \
        p.rc2++;                                                                                      \
    }

    if ((threadIdx.x >> 5) < V9_NWF) {
        // ================================================================================= geometry warps: flight
        v9_reg_dec<V9_RF>();
        unsigned n_cell = 0;
        long long idle0 = 0, wait0 = 0;
        const float Lux = S.Lux, Luy = S.Luy;
        const int ncx = S.ncx, ncy = S.ncy;
        const int cmx = (1 << S.shx) - 1, cmy = (1 << S.shy) - 1;
        const float* __restrict__ majp = S.maj;
        for (;;) {
            // every decision that rests on a volatile look at the control block is taken by lane 0 and broadcast: lanes
            // that have not reconverged would otherwise read different values and part ways around warp collectives
            const int navail = BCAST(v9_avail(ctl, RQ_FLY));
            if (navail < V9_MINB) {
                if (navail <= 0) {
                    if (BCAST(v9_ld_volatile(&ctl->retired)) >= unsigned(NPB)) break;
                    V9_IDLE();
                    continue;
                }
                idle0 = 0;
                const long long now = BCAST(clock64());
                if (wait0 == 0) wait0 = now;
                if (now - wait0 < V9_WAIT) { __nanosleep(100); continue; }
            }
            wait0 = 0; idle0 = 0;
            int slot;
            const int n = POP(rF, RQ_FLY, slot);
            if (n == 0) continue;
            const bool have = lane < n;
            Photon p;
            if (have) pool_load_flight<NPB, PL>(pool, slot, p);
            const bool frozen = FZ && (p.flags & FL_FROZEN);
            float dux = frozen ? 0.0f : p.d.x * S.inv_Sx, duy = frozen ? 0.0f : p.d.y * S.inv_Sy, dzg = p.d.z;
            if (fabsf(dux) < 1e-20f) dux = 1e-20f;
            if (fabsf(duy) < 1e-20f) duy = 1e-20f;
            if (fabsf(dzg) < 1e-12f) dzg = 1e-12f;
            const float kx = 1.0f / dux, ky = 1.0f / duy, kz = 1.0f / dzg;
            const bool upz = dzg > 0.0f;
            const int ox = dux > 0.0f ? 1 : 0, oy = duy > 0.0f ? 1 : 0;
            const int upmx = -ox, upmy = -oy;
            float ux = p.x, uy = p.y;
            int ev = EV_NONE;
#pragma unroll 1
            for (int kstep = 0; kstep < S.flight_steps; ++kstep) {
                if (have && ev == EV_NONE) {
                    if (UZ && !PL && (p.flags & FL_STALE)) {
                        p.is = S.uz_s0 + min(S.ncz - 1, max(0, __float2int_rd((p.z - S.uz_z0) * S.uz_inv)));
                        p.flags &= ~FL_STALE;
                    }
                    float4 A = sm.slabA[p.is];
                    int aw = __float_as_int(A.w);
                    const bool in3 = aw >= 0;
                    float mj = -1.0f;
                    if (in3) { mj = __ldg(majp + ((aw & 0xffff) * ncy + p.ciy) * ncx + p.cix); ++n_cell; }
                    if (!UZ && !PL && mj >= 0.0f && (p.flags & FL_STALE)) {
                        const int gw = __float_as_int(sm.grpA[(aw >> 16) & 0x7fff].w);
                        int lo = gw & 0xffff, hi = int(unsigned(gw) >> 16) - 1;
                        while (lo < hi) {
                            const int mid = (lo + hi + 1) >> 1;
                            if (p.z >= sm.slabA[mid].x) lo = mid; else hi = mid - 1;
                        }
                        p.is = lo;
                        A = sm.slabA[lo]; aw = __float_as_int(A.w);
                        mj = fmaxf(0.0f, __ldg(majp + ((aw & 0xffff) * ncy + p.ciy) * ncx + p.cix));
                    }
                    const bool empty = mj < 0.0f;
                    const int grp = (aw >> 16) & 0x7fff;
                    const int code = (in3 && empty) ? __float2int_rn(-mj) - 1 : (grp | (grp << 12));
                    const int glo = code & 0xfff, ghi = (code >> 12) & 0xfff;
                    const float4 G = sm.grpA[glo], Gh = sm.grpA[ghi];
                    const int mx = in3 ? (empty ? cmx : 0) : 0x3fffffff, my = in3 ? (empty ? cmy : 0) : 0x3fffffff;
                    const int bxlo = p.cix & ~mx, bylo = p.ciy & ~my;
                    const int fxi = bxlo + ((mx + 1) & upmx), fyi = bylo + ((my + 1) & upmy);
                    const float fxf = fminf(float(fxi), Lux), fyf = fminf(float(fyi), Luy);
                    const float zf = upz ? (empty ? Gh.y : A.y) : (empty ? G.x : A.x);
                    const float tx = (fxf - ux) * kx, ty = (fyf - uy) * ky, tz = (zf - p.z) * kz;
                    const float M = empty ? (glo == ghi ? G.z : S.maj1d_blk) : A.z + mj;
                    const float dexit = fmaxf(0.0f, fminf(tz, fminf(tx, ty)));
                    const bool hit = p.tau < M * dexit;
                    const float dmove = hit ? __fdividef(p.tau, M) : dexit;
                    const bool zc = !hit && (tz <= tx) && (tz <= ty);
                    const bool xc = !hit && !zc && (tx <= ty);
                    const bool yc = !hit && !zc && !xc;

                    if (PL && (p.flags & FL_ABS)) {
                        const float wn = p.w * __expf(-__ldg(S.job_abs + size_t(p.job) * S.nz + p.is) * dmove);
                        ACC_ADD(ACC_ATM, double(p.w) - double(wn));
                        if (want_heat) {
                            p.x = ux * S.Sx; p.y = uy * S.Sy;
                            heat_tally<false>(S, sm, p, p.is, double(p.w) - double(wn));
                        }
                        p.w = wn;
                    }
                    p.leg += dmove;
                    p.z = zc ? zf : p.z + dzg * dmove;
                    ux += dux * dmove; uy += duy * dmove;
                    p.tau = fmaxf(0.0f, p.tau - M * dmove);

                    int cxn = fxi - 1 + ox, cyn = fyi - 1 + oy;
                    float uxn = fxf, uyn = fyf;
                    if (fxi >= ncx) { cxn = 0; uxn = 0.0f; }
                    if (cxn < 0) { cxn = ncx - 1; uxn = Lux; }
                    if (fyi >= ncy) { cyn = 0; uyn = 0.0f; }
                    if (cyn < 0) { cyn = ncy - 1; uyn = Luy; }
                    const int cxi = min(min(p.cix | mx, ncx - 1), max(bxlo, __float2int_rd(ux)));
                    const int cyi = min(min(p.ciy | my, ncy - 1), max(bylo, __float2int_rd(uy)));
                    p.cix = xc ? cxn : cxi; ux = xc ? uxn : ux;
                    p.ciy = yc ? cyn : cyi; uy = yc ? uyn : uy;

                    const int slo = empty ? (__float_as_int(G.w) & 0xffff) : p.is;
                    const int shi = empty ? int(unsigned(__float_as_int(Gh.w)) >> 16) : p.is + 1;
                    int fl = p.flags;
                    if (mj >= 0.0f) fl &= ~FL_STALE;
                    if ((xc || yc) && in3 && shi - slo > 1) fl |= FL_STALE;
                    if (hit) {
                        ev = EV_TENT;
                        p.M = M;
                        fl = (fl & ~(FL_IN3 | FL_EMPTY)) | (in3 ? FL_IN3 : 0) | (empty ? FL_EMPTY : 0);
                    }
                    if (zc) {
                        fl &= ~FL_STALE;
                        if (PL && want_flux) {
                            p.x = ux * S.Sx; p.y = uy * S.Sy;
                            if (upz) flux_tally<false>(S, sm, p, 2, p.is + 1);
                            else {
                                if (p.flags & FL_DIRECT) flux_tally<false>(S, sm, p, 0, p.is);
                                flux_tally<false>(S, sm, p, 1, p.is);
                            }
                        }
                        const int nis = upz ? shi : slo - 1;
                        if (nis >= S.nslab_z) { ev = EV_ESC; fl |= FL_ESC; }
                        else if (nis < 0) ev = EV_SFC;
                        else p.is = nis;
                    }
                    p.flags = fl;
                }
                const unsigned flying = __ballot_sync(FULL, have && ev == EV_NONE);
                if (flying == 0u || n - __popc(flying) >= S.event_min) break;
            }
            if (have) {
                p.x = ux; p.y = uy;
                pool_store_flight<NPB, PL>(pool, slot, p);
            }
            v9_fence();
            PUSH(rF, RQ_FLY, have && ev == EV_NONE, slot);
            PUSH(rE, RQ_TENT, have && (ev == EV_TENT || ev == EV_ESC), slot);
            PUSH(rS, RQ_SFC, have && ev == EV_SFC, slot);
        }
        {
            unsigned long long v = n_cell;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
            if (lane == 0 && v) atomicAdd(reinterpret_cast<unsigned long long*>(S.stats) + 1, v);
        }
    } else {
        // ================================================================================= event warps
        v9_reg_inc<V9_RE>();
        const int nxy = S.nx * S.ny;
        long long idle0 = 0, wait0 = 0;
// a dead slot goes back to the DEAD ring, or is retired once the photon source is exhausted
#define KILL(cond)                                                                                   \
    {                                                                                                \
        if (BCAST(v9_ld_volatile(&ctl->exhausted))) {                                                \
            const unsigned m_ = __ballot_sync(FULL, (cond));                                         \
            if (lane == 0 && m_) atomicAdd(&ctl->retired, unsigned(__popc(m_)));                     \
        } else PUSH(rD, RQ_DEAD, (cond), slot);                                                      \
    }
        for (;;) {
            // ======================================================= pick the fullest ring
            int phase, nbest;
            {
                // lanes 0 ... 3 look at one ring each; everybody decides on the same four numbers (see the geometry warps)
                const int av = v9_avail(ctl, lane == 0 ? RQ_DEAD : (lane == 1 ? RQ_TENT : (lane == 2 ? RQ_COLL : RQ_SFC)));
                const int nD = __shfl_sync(FULL, av, 0), nE = __shfl_sync(FULL, av, 1), nC = __shfl_sync(FULL, av, 2), nS = __shfl_sync(FULL, av, 3);
                phase = 0; nbest = nD;
                if (nE >= nbest) { phase = 2; nbest = nE; }
                if (nC >= nbest) { phase = 3; nbest = nC; }
                if (nS > nbest) { phase = 4; nbest = nS; }
            }
            if (nbest < V9_MINB) {
                if (nbest <= 0) {
                    if (BCAST(v9_ld_volatile(&ctl->retired)) >= unsigned(NPB)) break;
                    V9_IDLE();
                    continue;
                }
                idle0 = 0;
                const long long now = BCAST(clock64());
                if (wait0 == 0) wait0 = now;
                if (now - wait0 < V9_WAIT) { __nanosleep(100); continue; }
            }
            wait0 = 0; idle0 = 0;

            Photon p;
            int slot;
            if (phase == 0) {
                // ======================================================= regeneration
                const int n = POP(rD, RQ_DEAD, slot);
                if (n == 0) continue;
                if (BCAST(v9_ld_volatile(&ctl->exhausted))) {
                    if (lane == 0) atomicAdd(&ctl->retired, unsigned(n));
                    continue;
                }
                const bool have = lane < n;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(S.counter, (unsigned long long)n);
                base = __shfl_sync(FULL, base, 0);
                const bool exhausted = base + (unsigned long long)n >= S.nphot_local;
                const unsigned long long idx = base + (unsigned long long)lane;
                const bool born = have && idx < S.nphot_local;
                if (born) {
                    int lo = 0, hi = S.njob - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (S.jobs[mid].first <= idx) lo = mid; else hi = mid - 1;
                    }
                    p.job = lo;
                    const DevJob& J = S.jobs[lo];
                    p.flags = FL_DIRECT | (J.has_abs ? FL_ABS : 0) | (J.has_fscale ? FL_FSCALE : 0);
                    const unsigned long long gidx = (unsigned long long)S.shard_rank + (idx - J.first) * (unsigned long long)S.shard_world;
                    p.rc0 = unsigned(gidx); p.rc1 = unsigned(gidx >> 32); p.rc2 = 0;
                    float4 u, v;
                    RNG4(u);
                    RNG4(v);
                    p.x = u.x * S.Lx; p.y = u.y * S.Ly; p.z = sm.z[S.nz];
                    if (S.src_cos_half < 1.0f) p.d = rotate_dir(S.src, 1.0f - u.z * (1.0f - S.src_cos_half), RT_2PI * u.w);
                    else p.d = S.src;
                    p.w = 1.0f; p.order = 0;
                    p.is = S.nslab_z - 1; p.iz = S.nz - 1;
                    p.za = p.z; p.iza = p.iz; p.leg = 0.0f; p.M = 0.0f;
                    p.cix = min(S.ncx - 1, int(p.x * S.inv_Sx));
                    p.ciy = min(S.ncy - 1, int(p.y * S.inv_Sy));
                    if (FZ && S.solver == B200RT_SOLVER_IPA) {
                        p.flags |= FL_FROZEN;
                        p.x = (float(p.cix) + 0.5f) * S.dx; p.y = (float(p.ciy) + 0.5f) * S.dy;
                    }
                    p.tau = -__logf(v.x);
                    CNT_ADD(CNT_PHOT, 1u);
                    if (want_flux) { flux_tally<false>(S, sm, p, 0, S.nz); flux_tally<false>(S, sm, p, 1, S.nz); }
                    pool_store<NPB>(pool, slot, p, S.inv_Sx, S.inv_Sy);
                }
                v9_fence();
                PUSH(rF, RQ_FLY, born, slot);
                if (exhausted) {
                    const unsigned m_ = __ballot_sync(FULL, have && !born);
                    if (lane == 0) {
                        *reinterpret_cast<volatile unsigned*>(&ctl->exhausted) = 1u;
                        if (m_) atomicAdd(&ctl->retired, unsigned(__popc(m_)));
                    }
                }
                continue;
            }

            if (phase == 2) {
                // ======================================================= tentative collisions (and escapes)
                const int n = POP(rE, RQ_TENT, slot);
                if (n == 0) continue;
                const bool have = lane < n;
                int ev = EV_NONE;
                if (have) {
                    pool_load<NPB>(pool, slot, p, S.Sx, S.Sy);
                    ev = (p.flags & FL_ESC) ? EV_ESC : EV_TENT;
                }
                bool accepted = false, rejected = false;
                float c_apf = 0.0f, c_uz = 0.0f, c_uw = 0.0f, c_s3 = 0.0f;
                if (ev == EV_ESC) {
                    if (!PL && (p.flags & FL_ABS)) {
                        const float inv_absdz = p.d.z != 0.0f ? fabsf(1.0f / p.d.z) : RT_INF;
                        const float wn = p.w * __expf(-abs_tau(S, sm, p.job, p.za, p.iza, p.z, S.nz - 1, p.leg, inv_absdz));
                        ACC_ADD(ACC_ATM, double(p.w) - double(wn));
                        p.w = wn;
                    }
                    ACC_ADD(ACC_TOA, double(p.w));
                } else if (ev == EV_TENT) {
                    const bool frozen = FZ && (p.flags & FL_FROZEN);
                    const bool ev_empty = (p.flags & FL_EMPTY) != 0;
                    const bool ev_in3 = (p.flags & FL_IN3) != 0;
                    float4 u;
                    RNG4(u);
                    {
                        const bool by_height = !PL && UZ && ev_empty && ev_in3;
                        if (by_height) p.is = S.uz_s0 + min(S.ncz - 1, max(0, __float2int_rd((p.z - S.uz_z0) * S.uz_inv)));
                        const int4 sb = sm.slabB[p.is];
                        int l0 = sb.x, l1 = sb.y;
                        if (ev_empty && !by_height) { const int4 gb = sm.grpB[sb.z]; l0 = gb.z; l1 = gb.w; }
                        p.iz = (l1 - l0 > 1) ? find_layer(sm, l0, l1, p.z) : l0;
                    }
                    const int izn = p.iz;
                    float sig = sm.e1tot[izn];
                    float s3 = 0.0f;
                    int fx = 0, fy = 0, vox = 0;
                    if (ev_in3) {
                        if (frozen) { fx = p.cix; fy = p.ciy; }
                        else {
                            const int shx = ev_empty ? S.shx : 0, shy = ev_empty ? S.shy : 0;
                            const int ixlo = (p.cix >> shx) << shx, ixhi = ixlo + (1 << shx);
                            const int iylo = (p.ciy >> shy) << shy, iyhi = iylo + (1 << shy);
                            fx = min(min(S.nx, ixhi * S.svx) - 1, max(ixlo * S.svx, int(p.x * S.inv_dx)));
                            fy = min(min(S.ny, iyhi * S.svy) - 1, max(iylo * S.svy, int(p.y * S.inv_dy)));
                        }
                        vox = ((izn - S.iz0) * S.ny + fy) * S.nx + fx;
                        if (!ev_empty) { s3 = __ldg(S.ext3tot + vox); sig += s3; CNT_ADD(CNT_TENT, 1u); }
                    }
                    p.tau = -__logf(u.y);
                    float uc = u.x * p.M;
                    if (!(uc < sig)) rejected = true;
                    else {
                        if (!PL && (p.flags & FL_ABS)) {
                            const float inv_absdz = p.d.z != 0.0f ? fabsf(1.0f / p.d.z) : RT_INF;
                            const float wn = p.w * __expf(-abs_tau(S, sm, p.job, p.za, p.iza, p.z, izn, p.leg, inv_absdz));
                            ACC_ADD(ACC_ATM, double(p.w) - double(wn));
                            p.w = wn;
                        }
                        float omg = 1.0f, apf = 0.0f;
                        bool found = false;
                        if (uc < s3) {
                            if (S.np3d == 1) {
                                const float2 pr = __ldg(S.prop3 + vox);
                                omg = pr.x; apf = pr.y; found = true;
                            } else {
                                const size_t n3 = size_t(S.nz3) * nxy;
                                for (int k = 0; k < S.np3d; ++k) {
                                    const float e = __ldg(S.ext3 + size_t(k) * n3 + vox);
                                    if (uc < e || k == S.np3d - 1) {
                                        const float2 pr = __ldg(S.prop3 + size_t(k) * n3 + vox);
                                        omg = pr.x; apf = pr.y; found = true;
                                        break;
                                    }
                                    uc -= e;
                                }
                            }
                        } else uc -= s3;
                        if (!found) {
                            for (int k = 0; k < S.np1d; ++k) {
                                const float e = sm.e1[k * S.nz + izn];
                                if (uc < e || k == S.np1d - 1) { omg = sm.o1[k * S.nz + izn]; apf = sm.a1[k * S.nz + izn]; break; }
                                uc -= e;
                            }
                        }
                        CNT_ADD(CNT_COLL, 1u);
                        const float wn = p.w * omg;
                        if (wn < p.w) {
                            ACC_ADD(ACC_ATM, double(p.w) - double(wn));
                            if (want_heat) heat_tally<false>(S, sm, p, izn, double(p.w) - double(wn));
                        }
                        p.w = wn;
                        p.order++; p.flags &= ~FL_DIRECT;
                        if (p.w > 0.0f) {
                            accepted = true;
                            c_apf = apf; c_uz = u.z; c_uw = u.w; c_s3 = s3;
                            if (ev_in3 && !frozen) { p.cix = min(S.ncx - 1, fx / S.svx); p.ciy = min(S.ncy - 1, fy / S.svy); }
                        }
                    }
                }
                if (rejected) pool_store_reject<NPB>(pool, slot, p);
                if (accepted) pool_store_accept<NPB>(pool, slot, p, c_apf, c_uz, c_uw, c_s3);
                v9_fence();
                PUSH(rF, RQ_FLY, rejected, slot);
                PUSH(rC, RQ_COLL, accepted, slot);
                KILL(have && !rejected && !accepted);
                continue;
            }

            // =========================================================== collisions (phase 3) or surface hits (4)
            const int evk = phase == 3 ? EV_COLL : EV_SFC;
            int n;
            if (phase == 3) n = POP(rC, RQ_COLL, slot);
            else n = POP(rS, RQ_SFC, slot);
            if (n == 0) continue;
            const bool have = lane < n;
            float c_uw = 0.0f;
            if (have) { pool_load<NPB>(pool, slot, p, S.Sx, S.Sy); c_uw = pool[F_AUX * NPB + slot]; }
            bool alive = have;
            if (have) do {
                float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                float3 newd;
                float apf = 0.0f;
                int fx = 0, fy = 0;
                float s3 = 0.0f;
                int sfc_type = 0;
                float prm[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                const float3 wi = make_float3(-p.d.x, -p.d.y, -p.d.z);
                if (evk == EV_COLL) {
                    apf = p.leg; u.z = p.za; u.w = c_uw; s3 = p.M;
                    const bool ev_in3 = (p.flags & FL_IN3) != 0;
                    if (ev_in3) {
                        if (FZ && (p.flags & FL_FROZEN)) { fx = p.cix; fy = p.ciy; }
                        else {
                            fx = min(min(S.nx, (p.cix + 1) * S.svx) - 1, max(p.cix * S.svx, int(p.x * S.inv_dx)));
                            fy = min(min(S.ny, (p.ciy + 1) * S.svy) - 1, max(p.ciy * S.svy, int(p.y * S.inv_dy)));
                        }
                    }
                    if (FZ && S.solver == B200RT_SOLVER_PARTIAL_3D && p.order >= S.iso_ss && !(p.flags & FL_FROZEN)) {
                        if (!ev_in3) { p.cix = min(S.nx - 1, int(p.x * S.inv_dx)); p.ciy = min(S.ny - 1, int(p.y * S.inv_dy)); }
                        else { p.cix = fx; p.ciy = fy; }
                        p.x = (float(p.cix) + 0.5f) * S.dx; p.y = (float(p.ciy) + 0.5f) * S.dy;
                        p.flags |= FL_FROZEN;
                    }
                } else {
                    p.iz = 0; p.is = 0; p.z = sm.z[0]; p.flags &= ~FL_STALE;
                    if (!PL && (p.flags & FL_ABS)) {
                        const float inv_absdz = p.d.z != 0.0f ? fabsf(1.0f / p.d.z) : RT_INF;
                        const float wn = p.w * __expf(-abs_tau(S, sm, p.job, p.za, p.iza, p.z, 0, p.leg, inv_absdz));
                        ACC_ADD(ACC_ATM, double(p.w) - double(wn));
                        p.w = wn;
                    }
                    CNT_ADD(CNT_SFC, 1u);
                    RNG4(u);
                    const bool frozen = FZ && (p.flags & FL_FROZEN);
                    int sx, sy;
                    if (frozen) {
                        sx = min(S.sfc_nx - 1, int((float(p.cix) + 0.5f) / float(S.nx) * float(S.sfc_nx)));
                        sy = min(S.sfc_ny - 1, int((float(p.ciy) + 0.5f) / float(S.ny) * float(S.sfc_ny)));
                    } else {
                        sx = min(S.sfc_nx - 1, max(0, int(p.x * S.inv_Lx * float(S.sfc_nx))));
                        sy = min(S.sfc_ny - 1, max(0, int(p.y * S.inv_Ly * float(S.sfc_ny))));
                    }
                    const int sn = S.sfc_nx * S.sfc_ny, si = sy * S.sfc_nx + sx;
                    sfc_type = __ldg(S.sfc_type + si);
#pragma unroll
                    for (int q = 0; q < 5; ++q) prm[q] = __ldg(S.sfc_param + q * sn + si);
                    if (want_rad && S.nz3 > 0 && S.iz0 == 0) {
                        fx = frozen ? p.cix : min(S.nx - 1, max(0, int(p.x * S.inv_dx)));
                        fy = frozen ? p.ciy : min(S.ny - 1, max(0, int(p.y * S.inv_dy)));
                        s3 = __ldg(S.ext3tot + fy * S.nx + fx);
                    }
                }
                p.za = p.z; p.iza = p.iz; p.leg = 0.0f;

                if (want_rad) {
                    for (int k = 0; k < S.nrad; ++k) {
                        const DevSensor& se = S.sens[k];
                        if (CAM && se.kind == 1) {
                            int pix = 0;
                            unsigned nv = 0;
                            const bool in3 = (S.nz3 > 0) && p.iz >= S.iz0 && p.iz < S.iz0 + S.nz3;
                            const float c = camera_le(S, sm.z, sm.e1tot, sm.e1cum, se, p.x, p.y, p.z, p.iz, p.job, p.flags & FL_ABS, fx, fy, in3, p.d, evk, apf,
                                                      sfc_type, prm[0], prm[1], prm[2], prm[3], prm[4], &pix, &nv);
                            CNT_ADD(CNT_VISIT, nv);
                            if (c > 0.0f) {
                                const DevJob& J = S.jobs[p.job];
                                tally_add(S.rad + size_t(J.slab) * S.rad_slab + se.off + pix, double(c * p.w) * J.rad_fac * se.npix);
                                CNT_ADD(CNT_LE, 1u);
                                CNT_ADD(CNT_TALLY, 1u);
                            }
                            continue;
                        }
                        const float dzs = (se.zt - p.z) * se.s.z;
                        if (!(dzs > 0.0f)) continue;
                        float f;
                        if (evk == EV_COLL) {
                            const float cosang = p.d.x * se.s.x + p.d.y * se.s.y + p.d.z * se.s.z;
                            f = phase_eval(S.pt, apf, cosang) * (0.25f / RT_PI);
                        } else {
                            f = se.s.z > 0.0f ? brdf_eval(sfc_type, prm[0], prm[1], prm[2], prm[3], prm[4], wi, se.s) * se.s.z : 0.0f;
                        }
                        if (f > 0.0f) le_deposit<false>(S, sm, se, p, f * p.w, fx, fy, s3);
                    }
                }

                if (evk == EV_COLL) {
                    float xi_tab = 0.5f;
                    if (apf >= 1.0f) { float4 v; RNG4(v); xi_tab = v.x; }
                    const float mu = phase_sample(S.pt, apf, u.z, xi_tab);
                    newd = rotate_dir(p.d, mu, RT_2PI * u.w);
                    if (p.order >= S.iso_max) { ACC_ADD(ACC_RR, -(double(p.w))); alive = false; break; }
                } else {
                    float3 wo;
                    const float fac = surface_sample(sfc_type, prm[0], prm[1], prm[2], prm[3], prm[4], wi, u, &wo);
                    const float wn = p.w * fac;
                    ACC_ADD(ACC_SFC, double(p.w) - double(wn));
                    p.w = wn;
                    if (!(p.w > 0.0f)) { alive = false; break; }
                    const float nrm = rsqrtf(wo.x * wo.x + wo.y * wo.y + wo.z * wo.z);
                    newd = make_float3(wo.x * nrm, wo.y * nrm, wo.z * nrm);
                    p.flags &= ~FL_DIRECT; p.order++;
                }
                p.d = newd;
                if (evk == EV_SFC) {
                    if (want_flux) flux_tally<false>(S, sm, p, 2, 0);
                    if (S.nz3 > 0 && S.iz0 == 0 && !(FZ && (p.flags & FL_FROZEN))) {
                        p.cix = min(S.ncx - 1, max(0, int(p.x * S.inv_Sx)));
                        p.ciy = min(S.ncy - 1, max(0, int(p.y * S.inv_Sy)));
                    }
                    if (FZ && S.solver == B200RT_SOLVER_PARTIAL_3D && p.order >= S.iso_ss && !(p.flags & FL_FROZEN)) {
                        p.cix = min(S.nx - 1, max(0, int(p.x * S.inv_dx))); p.ciy = min(S.ny - 1, max(0, int(p.y * S.inv_dy)));
                        p.x = (float(p.cix) + 0.5f) * S.dx; p.y = (float(p.ciy) + 0.5f) * S.dy;
                        p.flags |= FL_FROZEN;
                    }
                }
                if (p.w < S.wmin) {
                    float xi = u.w;
                    if (evk == EV_COLL) { float4 v; RNG4(v); xi = v.x; }
                    if (xi * S.wfac < p.w) { ACC_ADD(ACC_RR, double(S.wfac) - double(p.w)); p.w = S.wfac; }
                    else { ACC_ADD(ACC_RR, -(double(p.w))); CNT_ADD(CNT_KILL, 1u); alive = false; break; }
                }
                if (p.w < 1e-30f) { ACC_ADD(ACC_RR, -(double(p.w))); alive = false; break; }
            } while (0);

            if (have && alive) pool_store<NPB>(pool, slot, p, S.inv_Sx, S.inv_Sy);
            v9_fence();
            PUSH(rF, RQ_FLY, have && alive, slot);
            KILL(have && !alive);
        }
#undef KILL
    }
#undef RNG4
#undef PUSH
#undef POP
#undef V9_IDLE
#undef BCAST

    // ---- flush (every warp of the block has left its role loop)
    __syncthreads();
    if (PL && (sm.ftal || sm.htal)) {
        for (int i = threadIdx.x; i < S.ntal_flux_smem; i += blockDim.x) {
            const double v = sm.ftal[i];
            if (v != 0.0) { tally_add(S.flux + i, v); CNT_ADD(CNT_TALLY, 1u); }
        }
        for (int i = threadIdx.x; i < S.ntal_heat_smem; i += blockDim.x) {
            const double v = sm.htal[i];
            if (v != 0.0) { tally_add(S.heat + i, v); CNT_ADD(CNT_TALLY, 1u); }
        }
    }
    {
        double a = sm.acc_atm[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(FULL, a, o);
        if (lane == 0 && a != 0.0) atomicAdd(&S.stats->w_atm, a);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int slot_of[8] = {0, 2, 3, 4, 5, 6, 7, 8};
        unsigned long long* sc = reinterpret_cast<unsigned long long*>(S.stats);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned long long v = sm.cnt[i * 32 + lane];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
            if (lane == 0 && v) atomicAdd(sc + slot_of[i], v);
        }
        double dsum[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double v = sm.acc[i * 32 + lane];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
            dsum[i] = v;
        }
        if (lane == 0) {
            atomicAdd(&S.stats->w_toa, dsum[0]); atomicAdd(&S.stats->w_sfc, dsum[1]);
            atomicAdd(&S.stats->w_rr, dsum[3]);
        }
    }
}

typedef void (*transport_fn9)(const DevScene);
template <int NPB>
static transport_fn9 pick_v9_np(bool pl, bool fz, bool cam, bool uz) {
    if (cam) {
        if (pl) return uz ? transport_v9<true, false, NPB, true, true> : transport_v9<true, false, NPB, true, false>;
        return uz ? transport_v9<false, false, NPB, true, true> : transport_v9<false, false, NPB, true, false>;
    }
    if (pl) {
        if (fz) return uz ? transport_v9<true, true, NPB, false, true> : transport_v9<true, true, NPB, false, false>;
        return uz ? transport_v9<true, false, NPB, false, true> : transport_v9<true, false, NPB, false, false>;
    }
    if (fz) return uz ? transport_v9<false, true, NPB, false, true> : transport_v9<false, true, NPB, false, false>;
    return uz ? transport_v9<false, false, NPB, false, true> : transport_v9<false, false, NPB, false, false>;
}
static transport_fn9 pick_v9(bool pl, bool fz, bool cam, bool uz, int npb) {
    switch (npb) {
#ifdef V9_ALL_POOLS
        case 1024: return pick_v9_np<1024>(pl, fz, cam, uz);
        case 2048: return pick_v9_np<2048>(pl, fz, cam, uz);
#endif
        default: return pick_v9_np<V9_NPB>(pl, fz, cam, uz);
    }
}
