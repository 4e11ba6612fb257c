// sf_sor_stream.cu -- K4, variant 2: red-black SOR as a STREAMING wavefront (sor_coupled, solver.c:63-399, re-ordered).
//
// The arithmetic per pixel and half sweep is that of sf_sor.cu (same FMA chain, same colour order); what changes is the
// traversal.  The tiled kernel recomputes a 2T-pixel halo on all four sides of a 64x64 register tile (44 % of its
// relaxations at T = 4) and loads / relaxes / stores a tile one phase after the other.  Here a CTA owns a strip of 256
// columns and streams down a segment of rows; T sweeps = 2T half sweeps ("levels") are applied as a wavefront:
//
//     stream row j is loaded in step j - 1, relaxed at level k in step j + 1 + 2k, stored in step j + 17.
//
// Level k of row j needs rows j-1, j+1 after level k-1 and before level k+1 -- with a lag of two rows per level both
// hold, so all levels run at the same time on different rows, the vertical halo disappears (only the 8 rows above and
// below a SEGMENT are redundant) and loads, relaxations and stores of different rows overlap in time by construction.
//
//   * 9 warps; warp w owns the rows j = 2w, 2w+1 (mod 18) in REGISTERS: lane l holds columns 8l .. 8l+7 of both rows (7
//     coefficient planes + du,dv, as packed-fp32 pairs (column i, column i+4)).  In every step 8 warps relax one row each
//     (one colour: 4 pixels per lane = two FFMA2 chains) and the ninth stores a finished row and loads a new one, so a row
//     lives 18 steps.  Horizontal neighbours are in the lane except one column per side (one 32-bit shuffle each).
//   * Vertical neighbours: one of the two adjacent rows is the warp's other slot (registers); the other belongs to the
//     neighbouring warp and is read from a shared-memory exchange area that every relaxation publishes to.  Areas are
//     split by column parity and double-buffered by version, so the only ordering needed is "the neighbour has finished
//     step s-1 (slot 1) / s-3 (slot 0)": per-warp step counters in shared memory, no CTA barrier anywhere in the loop.
//     tools/proto/sor_stream_emu.py runs this protocol under a random scheduler against plain red-black sweeps.
//   * Rows arrive by TMA (cp.async.bulk.tensor.3d, a ring of SS_D row entries: 7 coefficient rows + du,dv + psi_v of the
//     row above); the warp that consumes an entry re-arms it for the row SS_D further down.  Out-of-image texels are
//     zero-filled, which gives the psi = 0 / a' = 0 boundary of the tiled kernel.
//   * Horizontal halo: 8 columns per side = lanes 0 and 31, which never store.
#include "sf_internal.cuh"
#include "sf_pack.cuh"
#include "sf_tma.cuh"

namespace sf {

constexpr int SS_NW = 9;                  // warps
constexpr int SS_WIN = 2 * SS_NW;         // stream rows in flight
constexpr int SS_SW = 256;                // strip width (8 columns per lane)
constexpr int SS_HALO = 8;                // redundant columns per side / rows above and below a segment (>= 2T)
constexpr int SS_IW = SS_SW - 2 * SS_HALO;
#ifndef SF_SS_D
#define SF_SS_D 12
#endif
constexpr int SS_D = SF_SS_D;             // staging ring depth in rows
constexpr int SS_ENTRY_FLOATS = 10 * SS_SW; // 7 coefficient rows, du, dv, psi_v of the row above
constexpr int SS_ENTRY_BYTES = SS_ENTRY_FLOATS * 4;
// exchange area: [row residue 18][buffer 2][column parity 2][du,dv][pair 2][lane 32] packed pairs
constexpr int SS_AREA_P64 = 2 * 2 * 32;                     // one (row, buffer, parity) area
constexpr int SS_EXCH_P64 = SS_WIN * 2 * 2 * SS_AREA_P64;
constexpr int SS_TAB_P64 = 10 * 32;                        // per-warp table: psi_v above slot 0 [4], phl of both slots [2], psi_v of slot 1 [4]
constexpr int SS_SMEM_BYTES = SS_D * SS_ENTRY_BYTES + SS_EXCH_P64 * 8 + SS_NW * SS_TAB_P64 * 8 + 64 /*step counters*/ +
                              SS_D * 8 /*mbarriers*/ + 1024;

#ifdef SF_SS_CLOCKS
__device__ unsigned long long g_ss_clk[8]; // flag wait, relax, publish-flag, reload total, of which TMA wait, -, total, warps
#define SS_CLK(var) const long long var = clock64()
#define SS_ACC(slot, t0, t1) clk_acc[slot] += (t1) - (t0)
#else
#define SS_CLK(var)
#define SS_ACC(slot, t0, t1)
#endif

#ifndef SF_SS_FENCED
#define SF_SS_FENCED 0 // 1: st.release / ld.acquire flags (a MEMBAR.ALL.CTA per publication); 0: volatile flags
#endif
#if SF_SS_FENCED
#define SS_ST_FLAG st_release_shared
#define SS_LD_FLAG ld_acquire_shared
#else
#define SS_ST_FLAG st_flag_shared
#define SS_LD_FLAG ld_flag_shared
#endif

struct SorStreamArgs {
    Geom g;
    float *out_du, *out_dv;
    int in_du_plane; // (du, dv) plane pair of the input iterate
    int nlev;        // half sweeps in this launch (2T <= 8)
    int strips, seg_rows;
    float omega;
    int zero_init;
    float one;       // 1.0f, opaque to the compiler (pairs are defined by a packed instruction, see load_row)
};

// one stream row in registers: packed pairs (column i, column i + 4), i = 0..3
struct SsRow {
    p64 na11[4], na12[4], na22[4]; // NEGATED inverse blocks
    p64 b1[4], b2[4];
    p64 ph[4], pv[4];
    p64 du[4], dv[4];
};

// One relaxation turn = the pixels of column parity XP of row `me` (slot B of the warp), in two parts.
// ss_prepare: everything that does not depend on the neighbouring warp -- the warp's own table entries (psi_v of the row
//   above slot 0 / of slot 1 itself, psi_h at the lane's left edge) and the two edge columns that come from the
//   neighbouring lanes -- is fetched BEFORE the warp polls its neighbour's step counter.
// ss_finish: the neighbouring warp's row (exchange area `nb`: this parity, the version of the previous level), the FMA
//   chains and the publication.  oth = the warp's other slot: the row below for slot 0, the row above for slot 1.
// Same chain as sor_relax_pair (sf_sor.cu): B = psl*dl + (psr*dr + (pst*dt + (psb*db + b))); n = (-a1)*B1 + ((-a2)*B2 + d);
// d += (-omega)*n.
struct SsPrep {
    p64 pa[2]; // psi_v of the row above (slot 0)
    p64 pb[2]; // psi_v of this row (slot 1: from the table)
    p64 phl;   // psi_h at columns (-1, 3)
    p64 eu, ev; // the pair that needs a neighbouring lane: (col -1, col 3) for parity 0, (col 4, col 8) for parity 1
};
template <int B, int XP>
__device__ __forceinline__ void ss_prepare(const SsRow &me, const p64 *__restrict__ tab, SsPrep &q) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (B == 0) q.pa[h] = tab[(XP + 2 * h) * 32];
        else q.pb[h] = tab[(6 + XP + 2 * h) * 32];
    }
    if (XP == 0) {
        q.phl = tab[(4 + B) * 32];
        q.eu = pk(__shfl_up_sync(0xffffffffu, hi_of(me.du[3]), 1), lo_of(me.du[3]));
        q.ev = pk(__shfl_up_sync(0xffffffffu, hi_of(me.dv[3]), 1), lo_of(me.dv[3]));
    } else {
        q.eu = pk(hi_of(me.du[0]), __shfl_down_sync(0xffffffffu, lo_of(me.du[0]), 1));
        q.ev = pk(hi_of(me.dv[0]), __shfl_down_sync(0xffffffffu, lo_of(me.dv[0]), 1));
    }
}
template <int B, int XP>
__device__ __forceinline__ void ss_finish(SsRow &me, const SsRow &oth, const SsPrep &q, const p64 *__restrict__ nb,
                                          p64 *__restrict__ mine, const p64 nomega2) {
    p64 su[2], sv[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        su[h] = nb[h * 32];
        sv[h] = nb[(2 + h) * 32];
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = XP + 2 * h; // compile-time after unrolling
        p64 ul, vl, ur, vr, psl;
        if (i == 0) {
            ul = q.eu; vl = q.ev; psl = q.phl;
        } else {
            ul = me.du[(i + 3) & 3]; vl = me.dv[(i + 3) & 3]; psl = me.ph[(i + 3) & 3];
        }
        if (i == 3) {
            ur = q.eu; vr = q.ev;
        } else {
            ur = me.du[(i + 1) & 3]; vr = me.dv[(i + 1) & 3];
        }
        const p64 psr = me.ph[i];
        p64 ut, vt, ub, vb, pst, psb;
        if (B == 0) {
            ut = su[h]; vt = sv[h]; ub = oth.du[i]; vb = oth.dv[i]; pst = q.pa[h]; psb = me.pv[i];
        } else {
            ut = oth.du[i]; vt = oth.dv[i]; ub = su[h]; vb = sv[h]; pst = oth.pv[i]; psb = q.pb[h];
        }
        const p64 B1 = fma2(psl, ul, fma2(psr, ur, fma2(pst, ut, fma2(psb, ub, me.b1[i]))));
        const p64 B2 = fma2(psl, vl, fma2(psr, vr, fma2(pst, vt, fma2(psb, vb, me.b2[i]))));
        const p64 nu = fma2(me.na11[i], B1, fma2(me.na12[i], B2, me.du[i]));
        const p64 nv = fma2(me.na12[i], B1, fma2(me.na22[i], B2, me.dv[i]));
        me.du[i] = fma2(nomega2, nu, me.du[i]);
        me.dv[i] = fma2(nomega2, nv, me.dv[i]);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        mine[h * 32] = me.du[XP + 2 * h];
        mine[(2 + h) * 32] = me.dv[XP + 2 * h];
    }
}

__global__ void __launch_bounds__(SS_NW * 32, 1)
k_sor_stream(const __grid_constant__ CUtensorMap tmap_coef, const __grid_constant__ CUtensorMap tmap_iter,
             const __grid_constant__ CUtensorMap tmap_row, SorStreamArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // (the shuffle makes the warp index provably warp-uniform: everything derived from it can live in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    float *ring = reinterpret_cast<float *>(base);
    p64 *exch = reinterpret_cast<p64 *>(base + SS_D * SS_ENTRY_BYTES);
    p64 *pvt_all = exch + SS_EXCH_P64;
    int *cnt = reinterpret_cast<int *>(pvt_all + SS_NW * SS_TAB_P64);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(cnt + 16);

    // ---- set-up that touches no global memory (may overlap the previous launch of the chain)
    for (int k = threadIdx.x; k < SS_EXCH_P64 + SS_NW * SS_TAB_P64; k += SS_NW * 32) exch[k] = 0ull;
    // a warp has trivially "finished" every step before its first row arrives (row 2w is loaded in step 2w - 1)
    if (threadIdx.x < SS_NW) cnt[threadIdx.x] = 2 * (int)threadIdx.x - 2;
    if (threadIdx.x < SS_D) mbar_init(mbar + threadIdx.x, 1);
    if (threadIdx.x == 0) {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    pdl_enter();
    if (a.g.cancelled()) return;

    const int strip = blockIdx.x % a.strips, seg = blockIdx.x / a.strips;
    const int x0 = strip * SS_IW - SS_HALO;
    const int ya = seg * a.seg_rows, yb = min(ya + a.seg_rows, a.g.H);
    const int ystart = ya - SS_HALO;          // even: seg_rows is even
    const int nrows = (yb - ya) + 2 * SS_HALO;
    const int s_last = nrows + 16;            // a slot changes rows for the last time in step nrows + 16
    // a row is stored right after its last level when that level closes a group of four (see the turn loop)
    const bool early_store = (a.nlev & 3) == 0;

    auto issue = [&](int j) { // one lane: pull stream row j into its ring entry
        const int e = j % SS_D;
        float *dst = ring + e * SS_ENTRY_FLOATS;
        uint64_t *bar = mbar + e;
        const int y = ystart + j;
        mbar_expect_tx(bar, (uint32_t)(((a.zero_init ? 7 : 9) + 1) * SS_SW * 4));
        tma_load_3d(dst, &tmap_coef, bar, x0, y, 0);
        if (!a.zero_init) tma_load_3d(dst + 7 * SS_SW, &tmap_iter, bar, x0, y, a.in_du_plane);
        tma_load_3d(dst + 9 * SS_SW, &tmap_row, bar, x0, y - 1, SP_PV);
    };
    if (threadIdx.x == 0) {
        const int n0 = nrows < SS_D ? nrows : SS_D;
        for (int j = 0; j < n0; j++) issue(j);
    }

    const p64 z2 = pk(0.0f, 0.0f);
    const p64 nomega2 = pk(-a.omega, -a.omega);
    SsRow R0, R1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        R0.na11[i] = R0.na12[i] = R0.na22[i] = R0.b1[i] = R0.b2[i] = R0.ph[i] = R0.pv[i] = R0.du[i] = R0.dv[i] = z2;
        R1.na11[i] = R1.na12[i] = R1.na22[i] = R1.b1[i] = R1.b2[i] = R1.ph[i] = R1.pv[i] = R1.du[i] = R1.dv[i] = z2;
    }

    // exchange areas: [row residue][buffer][column parity] -> 128 packed pairs; this warp's rows are the residues 2w and
    // 2w + 1, the rows next to them 2w - 1 and 2w + 2 (modulo 18)
    constexpr int RES = 2 * 2 * SS_AREA_P64, BUF = 2 * SS_AREA_P64, PAR = SS_AREA_P64;
    p64 *const ex0 = exch + lane + (2 * warp) * RES;
    p64 *const ex1 = ex0 + RES;
    const p64 *const ex_up = exch + lane + ((2 * warp + SS_WIN - 1) % SS_WIN) * RES;
    const p64 *const ex_dn = exch + lane + ((2 * warp + 2) % SS_WIN) * RES;
    p64 *const tab = pvt_all + warp * SS_TAB_P64 + lane; // per-warp table (SS_TAB_P64)
    const int *cnt_up = cnt + (warp + SS_NW - 1) % SS_NW, *cnt_dn = cnt + (warp + 1) % SS_NW;
    int *cnt_me = cnt + warp;
    // owner of stream row j - SS_D when this warp owns row j (row residues are 2w, 2w+1 modulo 18)
    const int *cnt_arm = cnt + (warp + SS_NW - (SS_D / 2) % SS_NW) % SS_NW;
    static_assert(SS_D % 2 == 0, "a ring entry must be handed between same-parity rows");

#ifdef SF_SS_CLOCKS
    long long clk_acc[6] = {0, 0, 0, 0, 0, 0};
    const long long clk_begin = clock64();
#endif
    // Every lane reads the neighbour's step counter (one broadcast LDS) and the loop condition is a warp vote, i.e.
    // uniform by construction: the warp stays converged.  (With a per-lane or a lane-0 loop the lanes left the loop in
    // groups that the hardware did not merge again: ncu showed the relaxation running with 16-21 active threads and the
    // out-of-line divergent-shuffle path being taken.)
    auto wait_for = [&](const int *flag, int need) {
        SS_CLK(w0);
        uint32_t spins = 0;
        while (!__all_sync(0xffffffffu, SS_LD_FLAG(flag) >= need)) {
            if (++spins > (1u << 22)) __trap();
        }
        SS_CLK(w1);
        SS_ACC(0, w0, w1);
    };
    auto done_step = [&](int s) {
        SS_CLK(d0);
        __syncwarp();
        if (lane == 0) SS_ST_FLAG(cnt_me, s);
        SS_CLK(d1);
        SS_ACC(2, d0, d1);
    };

    // shared -> registers for one row: the two float4 of a plane row are transposed into four (i, i+4) pairs, each
    // defined by one FFMA2 (x*1 + 0, x*(-1) + 0 for the negated blocks) so that ptxas keeps it in an aligned pair
    auto load_row = [&](SsRow &r, const float *entry) {
        const p64 one2 = pk(a.one, a.one), mone2 = pk(-a.one, -a.one);
        const float4 *e4 = reinterpret_cast<const float4 *>(entry) + 2 * lane;
        auto ld4 = [&](int plane, p64 (&dst)[4], p64 scale) {
            const float4 q0 = e4[plane * (SS_SW / 4)], q1 = e4[plane * (SS_SW / 4) + 1];
            dst[0] = fma2(pk(q0.x, q1.x), scale, z2);
            dst[1] = fma2(pk(q0.y, q1.y), scale, z2);
            dst[2] = fma2(pk(q0.z, q1.z), scale, z2);
            dst[3] = fma2(pk(q0.w, q1.w), scale, z2);
        };
        if (a.zero_init) {
#pragma unroll
            for (int i = 0; i < 4; i++) r.du[i] = r.dv[i] = z2;
        } else {
            ld4(7, r.du, one2);
            ld4(8, r.dv, one2);
        }
        ld4(SP_A11, r.na11, mone2);
        ld4(SP_A12, r.na12, mone2);
        ld4(SP_A22, r.na22, mone2);
        ld4(SP_B1, r.b1, one2);
        ld4(SP_B2, r.b2, one2);
        ld4(SP_PH, r.ph, one2);
        ld4(SP_PV, r.pv, one2);
    };

    // finished row -> global (lanes 1..30: the outer 8 columns of the strip are halo).  The lane index is re-read from
    // the special register here: kept live across the turn loop it was spilled to local memory (ncu: ~200 clocks of
    // long-scoreboard stall per store on the reload).
    auto store_row = [&](const SsRow &r, int j) {
        unsigned ln;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(ln));
        const int y = ystart + j, gx = x0 + 8 * (int)ln;
        if (y < ya || y >= yb || ln == 0 || ln == 31 || gx >= a.g.W) return;
        const size_t o = (size_t)y * a.g.S + gx;
        *reinterpret_cast<float4 *>(a.out_du + o) = make_float4(lo_of(r.du[0]), lo_of(r.du[1]), lo_of(r.du[2]), lo_of(r.du[3]));
        *reinterpret_cast<float4 *>(a.out_dv + o) = make_float4(lo_of(r.dv[0]), lo_of(r.dv[1]), lo_of(r.dv[2]), lo_of(r.dv[3]));
        if (gx + 4 < a.g.W) {
            *reinterpret_cast<float4 *>(a.out_du + o + 4) = make_float4(hi_of(r.du[0]), hi_of(r.du[1]), hi_of(r.du[2]), hi_of(r.du[3]));
            *reinterpret_cast<float4 *>(a.out_dv + o + 4) = make_float4(hi_of(r.dv[0]), hi_of(r.dv[1]), hi_of(r.dv[2]), hi_of(r.dv[3]));
        }
    };

    // The TMA pull that re-arms a consumed ring entry is deferred to the start of the warp's next turn (it has SS_D
    // steps of slack), so that the reload publishes as early as possible.
    int pend_issue = -1;
    auto flush_issue = [&]() {
        if (pend_issue >= 0) {
            if (lane == 0) issue(pend_issue);
            pend_issue = -1;
        }
    };

    // the turn in which a slot changes rows (step s): (store stream row s - 17,) load row s + 1, publish it (version 0)
    auto reload = [&](auto slot_tag, SsRow &r, int s) {
        constexpr int B = decltype(slot_tag)::value;
        flush_issue();
        const int jo = s - 17, jn = s + 1;
        const bool have = jn < nrows;
        const int e = jn % SS_D;
        const uint32_t par = (uint32_t)((jn / SS_D) & 1);
        // The entry was re-armed for this row by the consumer of row jn - SS_D in step jn - SS_D - 1: without this gate
        // a warp that starts early would find the barrier in its previous phase and pass (try_wait on the parity of an
        // older phase succeeds at once).  The barrier is tested before the flag wait: the test has a long latency.
        if (have && jn >= SS_D) wait_for(cnt_arm, s - SS_D);
        bool ok = have ? mbar_try_wait(mbar + e, par) : true;
        if (B == 0) wait_for(cnt_up, s - 3);
        else wait_for(cnt_dn, s - 1);
        SS_CLK(r0);
        if (jo >= 0 && !early_store) store_row(r, jo);
        if (have) {
            {
                SS_CLK(m0);
                uint32_t spins = 0;
                while (!__all_sync(0xffffffffu, ok)) {
                    ok = mbar_try_wait(mbar + e, par);
                    if (++spins > (1u << 24)) __trap();
                }
                SS_CLK(m1);
                SS_ACC(4, m0, m1);
            }
            const float *entry = ring + e * SS_ENTRY_FLOATS;
            load_row(r, entry);
            // version 0 of both column parities (du,dv are loaded first)
            p64 *ar = (B == 0) ? ex0 : ex1;
#pragma unroll
            for (int par2 = 0; par2 < 2; par2++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    ar[par2 * PAR + h * 32] = r.du[par2 + 2 * h];
                    ar[par2 * PAR + (2 + h) * 32] = r.dv[par2 + 2 * h];
                }
            {   // psi_h at columns (-1, 3): the left neighbour's last column (0 at the strip edge) and the own column 3
                const float left = __shfl_up_sync(0xffffffffu, hi_of(r.ph[3]), 1);
                tab[(4 + B) * 32] = pk(lane == 0 ? 0.0f : left, lo_of(r.ph[3]));
            }
            if (B == 1) { // slot 1's own psi_v lives in the table (the register file holds 168 registers per thread at 9 warps)
#pragma unroll
                for (int i = 0; i < 4; i++) tab[(6 + i) * 32] = r.pv[i];
            }
            if (B == 0) { // psi_v of the row above: entry row 9 -> the table
                const float4 *e4 = reinterpret_cast<const float4 *>(entry) + 2 * lane + 9 * (SS_SW / 4);
                const float4 q0 = e4[0], q1 = e4[1];
                tab[0] = pk(q0.x, q1.x);
                tab[32] = pk(q0.y, q1.y);
                tab[64] = pk(q0.z, q1.z);
                tab[96] = pk(q0.w, q1.w);
            }
            if (jn + SS_D < nrows) pend_issue = jn + SS_D;
        }
        SS_CLK(r1);
        SS_ACC(3, r0, r1);
        done_step(s); // (the warp barrier in here also closes the reads of the ring entry before it is re-armed)
    };

    struct Slot0 { enum { value = 0 }; };
    struct Slot1 { enum { value = 1 }; };
    SsPrep q;
#define SS_TURN(SLOT, XP, ME, OTH, FLAG, NEED, NB, MINE, LEVEL, STEP)                           \
    {                                                                                           \
        const bool on_ = (SLOT ? v1 : v0) && (LEVEL) < a.nlev;                                  \
        flush_issue();                                                                          \
        if (on_) ss_prepare<SLOT, XP>(ME, tab, q);                                              \
        wait_for(FLAG, NEED);                                                                   \
        SS_CLK(x0_);                                                                            \
        if (on_) ss_finish<SLOT, XP>(ME, OTH, q, NB, MINE, nomega2);                            \
        SS_CLK(x1_);                                                                            \
        SS_ACC(1, x0_, x1_);                                                                    \
        done_step(STEP);                                                                        \
    }

    for (int c = 0;; c++) {
        const int j0 = SS_WIN * c + 2 * warp; // stream rows of this cycle: j0 (slot 0), j0 + 1 (slot 1)
        int s = j0 - 1;
        if (s > s_last) break;
        reload(Slot0{}, R0, s);
        reload(Slot1{}, R1, s + 1);
        s += 2;
        const bool v0 = j0 < nrows, v1 = j0 + 1 < nrows;
#pragma unroll 1
        for (int kq = 0; kq < 2; kq++, s += 8) {
            // Levels k = 4kq .. 4kq+3.  Column parity (k + row) & 1; a reader takes version ((k+1)>>1) & 1 of the
            // neighbour's area, a writer produces version ((k>>1)+1) & 1: (0,1) (1,1) (1,0) (0,0) for k mod 4 = 0..3.
            // Slot 0 waits for the warp above to have finished step s - 3, slot 1 for the warp below step s - 1.
            const int k = 4 * kq;
            SS_TURN(0, 0, R0, R1, cnt_up, s - 3, ex_up + 0 * BUF + 0 * PAR, ex0 + 1 * BUF + 0 * PAR, k, s);
            SS_TURN(1, 1, R1, R0, cnt_dn, s, ex_dn + 0 * BUF + 1 * PAR, ex1 + 1 * BUF + 1 * PAR, k, s + 1);
            SS_TURN(0, 1, R0, R1, cnt_up, s - 1, ex_up + 1 * BUF + 1 * PAR, ex0 + 1 * BUF + 1 * PAR, k + 1, s + 2);
            SS_TURN(1, 0, R1, R0, cnt_dn, s + 2, ex_dn + 1 * BUF + 0 * PAR, ex1 + 1 * BUF + 0 * PAR, k + 1, s + 3);
            SS_TURN(0, 0, R0, R1, cnt_up, s + 1, ex_up + 1 * BUF + 0 * PAR, ex0 + 0 * BUF + 0 * PAR, k + 2, s + 4);
            SS_TURN(1, 1, R1, R0, cnt_dn, s + 4, ex_dn + 1 * BUF + 1 * PAR, ex1 + 0 * BUF + 1 * PAR, k + 2, s + 5);
            SS_TURN(0, 1, R0, R1, cnt_up, s + 3, ex_up + 0 * BUF + 1 * PAR, ex0 + 0 * BUF + 1 * PAR, k + 3, s + 6);
            // a row whose last level was k + 3 goes out now (after the step counter: nobody waits for the store)
            if (early_store && v0 && k + 4 == a.nlev) store_row(R0, j0);
            SS_TURN(1, 0, R1, R0, cnt_dn, s + 6, ex_dn + 0 * BUF + 0 * PAR, ex1 + 0 * BUF + 0 * PAR, k + 3, s + 7);
            if (early_store && v1 && k + 4 == a.nlev) store_row(R1, j0 + 1);
        }
    }
#undef SS_TURN
#ifdef SF_SS_CLOCKS
    if (lane == 0) {
        for (int q2 = 0; q2 < 5; q2++) atomicAdd(&g_ss_clk[q2], (unsigned long long)clk_acc[q2]);
        atomicAdd(&g_ss_clk[6], (unsigned long long)(clock64() - clk_begin));
        atomicAdd(&g_ss_clk[7], 1ull);
    }
#endif
    __syncwarp();
    if (lane == 0) SS_ST_FLAG(cnt_me, 0x7fffffff);
}

bool sor_stream_device_init() {
    if (cudaFuncSetAttribute(k_sor_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_SMEM_BYTES) != cudaSuccess) {
        set_error("cudaFuncSetAttribute(k_sor_stream) failed");
        return false;
    }
    return true;
}

void sor_stream_boxes(unsigned boxes[3][3]) {
    const unsigned b[3][3] = {{SS_SW, 1, 7}, {SS_SW, 1, 2}, {SS_SW, 1, 1}};
    for (int m = 0; m < 3; m++)
        for (int d = 0; d < 3; d++) boxes[m][d] = b[m][d];
}

int launch_sor_stream(cudaStream_t st, SorPlan &plan, int iterations, float omega, int fuse, int *cur, bool zero_init) {
    const Geom g = plan.g;
    float *A = plan.arena;
    const size_t P = g.plane();
    if (fuse < 1) fuse = 1;
    if (fuse > SS_HALO / 2) fuse = SS_HALO / 2;
    // launches of (almost) equal depth: 30 sweeps at fuse 4 = 4,4,4,4,4,4,3,3 (a shallow last launch costs as much as a
    // full one: the pipeline is as long)
    const int nl = (iterations + fuse - 1) / fuse;
    const int base = iterations / nl, extra = iterations % nl;
    const int strips = (g.W + SS_IW - 1) / SS_IW;
    int segs = plan.num_sms / strips;
    if (segs < 1) segs = 1;
    int seg_rows = (g.H + segs - 1) / segs;
    if (seg_rows < 16) seg_rows = 16;
    seg_rows = (seg_rows + 1) & ~1;
    segs = (g.H + seg_rows - 1) / seg_rows;
    int launches = 0;
    for (int l = 0; l < nl; l++) {
        SorStreamArgs a;
        a.g = g;
        a.nlev = 2 * (base + (l < extra ? 1 : 0));
        a.strips = strips;
        a.seg_rows = seg_rows;
        a.omega = omega;
        a.one = 1.0f;
        a.zero_init = (zero_init && l == 0) ? 1 : 0;
        a.in_du_plane = *cur ? SP_DUB : SP_DUA;
        a.out_du = A + (size_t)(*cur ? SP_DUA : SP_DUB) * P;
        a.out_dv = A + (size_t)(*cur ? SP_DVA : SP_DVB) * P;
        if (launch_pdl(k_sor_stream, dim3(strips * segs), dim3(SS_NW * 32), (size_t)SS_SMEM_BYTES, st, plan.tmap_s_coef,
                       plan.tmap_s_iter, plan.tmap_s_row, a) != cudaSuccess) {
            set_error("cudaLaunchKernelEx(k_sor_stream) failed");
            return -1;
        }
        *cur ^= 1;
        launches++;
    }
    return launches;
}

} // namespace sf
