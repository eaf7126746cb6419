// lbm_collide4.cuh -- the collision of the four cells of a thread with PACKED fp32 arithmetic (sm_100a FADD2 / FMUL2 /
// FFMA2: one instruction, two IEEE-rounded fp32 operations on an aligned register pair).
//
// Included by lbm_kernels.cuh after update_cell() (the per-cell code with its guarded IEEE slow paths, which this
// file falls back to for operands outside its windows).
//
// Why: every step kernel of this library is bound by instruction issue before it is bound by HBM (strict flavour:
// ~1130 executed instructions per warp and four cells, ~650 of them the reference's own fp32 operations).  The
// packed instructions perform exactly the scalar operation on each half (round to nearest even, denormals kept: no
// .ftz), so the strict flavour stays bit-identical to SerialCode/d2q9-bgk.c:325-401,425-452 while the fp32 part
// of the stream takes half the issue slots.  Cells 0,1 of a thread form one pair, cells 2,3 the other; a 128-bit
// shared-memory or global load of four consecutive cells lands in two aligned pairs as it is.
//
// Structure (per pair): update_cell() guards each of its special sequences (division by rho, the
// constant divisions, the square root) with its own test and branch; ptxas schedules within basic blocks, so here the
// moments of the pair are formed first and tested against an "early" window; a thread with a cell outside it
// (never in a physical flow) takes update_cell()'s guarded code for the pair; everybody else runs the sequences
// unguarded in one block, and a "late" window test on the new populations decides whether the |u| of a cell has to be
// redone with the guarded code (the new populations themselves are exact by then).  Windows:
//   early: 2^-20 <= rho <= 2^20 and max(|m_x|, |m_y|) <= 2 rho  (so |u| <~ 2)
//          => the division sequence is exact (rho in [2^-40, 2^40], |m| < 2^40, see div2_rn), the constant divisions
//             are exact (|u| < 2^58), and everything after them is plain IEEE arithmetic: the new populations are exact;
//   late:  the same test on the moments of the NEW populations (4 rho' as the momentum bound)
//          => the division is exact again, u'^2 <= 2^100, so the square-root sequence is exact where u'^2 >= 2^-100
//             and the result is 0 below (a select, as in speed_from_sq).
// NaN operands fail both tests.  The fast flavour uses the same structure around its own arithmetic.
#pragma once


namespace lbm {

typedef unsigned long long f2; // two fp32 values: low word = first cell of the pair

__device__ __forceinline__ f2 pk(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 bc2(float c) { return pk(c, c); }
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// a * b whose result feeds a packed add / sub.  ptxas 12.9 contracts mul.rn.f32x2 followed by add/sub.rn.f32x2 into
// FFMA2 although both carry an explicit rounding modifier (it honours it for the scalar forms), with -fmad=false
// too, and it sees through fma(a, b, -0.0) and fma(p, 1.0, c) with literal constants -- fatal for the bit-exact
// flavour.  Here the product is formed as fma(a, b, z) with z = -0.0f read from constant memory, a value the
// compiler cannot know: RN(a*b + -0) is RN(a*b) bit for bit (a zero product keeps its sign: +0 + -0 = +0,
// -0 + -0 = -0), and an FFMA2 cannot be merged with the add that follows.  lbm_selftest_collide() compares
// collide4() with the scalar update_cell() bit for bit on the device, so a toolchain that behaves differently is
// caught by the tests.
__constant__ float c_negzero = -0.0f;
// -a on both halves (ptxas folds it into the operand modifier of the consuming packed instruction)
__device__ __forceinline__ f2 neg2(f2 a)
{
    float lo, hi;
    upk(a, lo, hi);
    return pk(-lo, -hi);
}

__device__ __forceinline__ f2 mul2_rounded(f2 a, f2 b) { return fma2(a, b, bc2(c_negzero)); }

// ---- NP pairs side by side ("vertical" code) --------------------------------------------------------------------
// Every operation below is written for NP pairs at once, so the instruction stream ptxas sees already interleaves NP
// independent dependency chains (it keeps close to source order, and the collision is one long chain: density sum ->
// division -> equilibrium -> relaxation -> density sum -> division -> square root).  With 16 warps per SM the kernels
// are bound by exactly that latency; one pair after the other measured 5-10 % slower than two side by side.
template <int NP>
struct PV {
    f2 v[NP];
};
#define LBM_PV_OP2(name, expr)                                                      \
    template <int NP>                                                               \
    __device__ __forceinline__ PV<NP> name(const PV<NP>& a, const PV<NP>& b)        \
    {                                                                               \
        PV<NP> r;                                                                   \
        _Pragma("unroll") for (int h = 0; h < NP; h++) r.v[h] = expr;               \
        return r;                                                                   \
    }
LBM_PV_OP2(vadd, add2(a.v[h], b.v[h]))
LBM_PV_OP2(vsub, sub2(a.v[h], b.v[h]))
LBM_PV_OP2(vmul, mul2(a.v[h], b.v[h]))
LBM_PV_OP2(vmul_rounded, mul2_rounded(a.v[h], b.v[h]))
#undef LBM_PV_OP2
template <int NP>
__device__ __forceinline__ PV<NP> vfma(const PV<NP>& a, const PV<NP>& b, const PV<NP>& c)
{
    PV<NP> r;
#pragma unroll
    for (int h = 0; h < NP; h++) r.v[h] = fma2(a.v[h], b.v[h], c.v[h]);
    return r;
}
template <int NP>
__device__ __forceinline__ PV<NP> vneg(const PV<NP>& a)
{
    PV<NP> r;
#pragma unroll
    for (int h = 0; h < NP; h++) r.v[h] = neg2(a.v[h]);
    return r;
}
template <int NP>
__device__ __forceinline__ PV<NP> vbc(float c)
{
    PV<NP> r;
#pragma unroll
    for (int h = 0; h < NP; h++) r.v[h] = bc2(c);
    return r;
}

// a1 / b and a2 / b: div2_rn()'s sequence without its operand test (the caller's window implies it)
template <int NP>
__device__ __forceinline__ void vdiv2(const PV<NP>& a1, const PV<NP>& a2, const PV<NP>& b, PV<NP>& q1, PV<NP>& q2)
{
    PV<NP> y0;
#pragma unroll
    for (int h = 0; h < NP; h++) {
        float b0, b1, y00, y01;
        upk(b.v[h], b0, b1);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y00) : "f"(b0));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y01) : "f"(b1));
        y0.v[h] = pk(y00, y01);
    }
    const PV<NP> nb = vneg(b);
    const PV<NP> e = vfma(nb, y0, vbc<NP>(1.f));
    const PV<NP> y = vfma(y0, e, y0);
    const PV<NP> p1 = vmul(a1, y), p2 = vmul(a2, y);
    q1 = vfma(y, vfma(nb, p1, a1), p1);
    q2 = vfma(y, vfma(nb, p2, a2), p2);
}
// div_const():  fma(-q, c, x) == fma(q, -c, x) (the product is exact either way)
template <int NP>
__device__ __forceinline__ PV<NP> vdiv_const(const PV<NP>& x, float c, float rc)
{
    const PV<NP> q = vmul(x, vbc<NP>(rc));
    const PV<NP> r = vfma(q, vbc<NP>(-c), x);
    return vfma(r, vbc<NP>(rc), q);
}
// speed_from_sq()'s sequence, a select instead of its range test; s[2h], s[2h+1] = the roots of pair h
template <int NP>
__device__ __forceinline__ void vspeed_from_sq(const PV<NP>& x, float (&s)[2 * NP])
{
    PV<NP> y;
#pragma unroll
    for (int h = 0; h < NP; h++) {
        float x0, x1, y0, y1;
        upk(x.v[h], x0, x1);
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x0));
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(x1));
        y.v[h] = pk(y0, y1);
    }
    const PV<NP> g = vmul(x, y), hh = vmul(y, vbc<NP>(0.5f));
    const PV<NP> r = vfma(vfma(vneg(g), g, x), hh, g);
#pragma unroll
    for (int h = 0; h < NP; h++) {
        float x0, x1, r0, r1;
        upk(x.v[h], x0, x1), upk(r.v[h], r0, r1);
        s[2 * h] = x0 >= 7.888609052210118e-31f /* 2^-100 */ ? r0 : 0.f;
        s[2 * h + 1] = x1 >= 7.888609052210118e-31f ? r1 : 0.f;
    }
}
__device__ __forceinline__ bool in_window(float rho, float mx, float my, float lo, float hi, float bound)
{
    return (rho >= lo) & (rho <= hi) & (fmaxf(fabsf(mx), fabsf(my)) <= __fmul_rn(bound, rho)); // no short circuit: no branches
}
// all cells of the NP pairs inside the window, obstacle cells (bit j of solid) excepted
template <int NP>
__device__ __forceinline__ bool all_in_window(const PV<NP>& rho, const PV<NP>& mx, const PV<NP>& my, uint32_t solid, float lo, float hi,
                                              float bound)
{
    bool ok = true;
#pragma unroll
    for (int h = 0; h < NP; h++) {
        float r0, r1, a0, a1, b0, b1;
        upk(rho.v[h], r0, r1), upk(mx.v[h], a0, a1), upk(my.v[h], b0, b1);
        ok = ok & (in_window(r0, a0, b0, lo, hi, bound) | (((solid >> (2 * h)) & 1u) != 0)) &
             (in_window(r1, a1, b1, lo, hi, bound) | (((solid >> (2 * h + 1)) & 1u) != 0));
    }
    return ok;
}

// |u| of a stored cell with the fast flavour's arithmetic, guarded (collide_cell<false>'s last lines)
__device__ __forceinline__ float speed_fast_guarded(const float c[Q])
{
    float r2 = c[0];
#pragma unroll
    for (int k = 1; k < Q; k++) r2 += c[k];
    const float nx_ = (c[1] + c[5] + c[8]) - (c[3] + c[6] + c[7]);
    const float ny_ = (c[2] + c[5] + c[6]) - (c[4] + c[7] + c[8]);
    return __fdividef(speed_from_sq(fmaf(nx_, nx_, ny_ * ny_)), r2);
}

// rho and the momentum; STRICT: SerialCode/d2q9-bgk.c:325-349 (sequential density sum from 0.f, velocity brackets
// left to right)
template <bool STRICT, int NP>
__device__ __forceinline__ void vmoments(const PV<NP> (&t)[Q], PV<NP>& rho, PV<NP>& mx, PV<NP>& my)
{
    PV<NP> d = STRICT ? vadd(vbc<NP>(0.f), t[0]) : t[0];
#pragma unroll
    for (int k = 1; k < Q; k++) d = vadd(d, t[k]);
    rho = d;
    mx = vsub(vadd(vadd(t[1], t[5]), t[8]), vadd(vadd(t[3], t[6]), t[7]));
    my = vsub(vadd(vadd(t[2], t[5]), t[6]), vadd(vadd(t[4], t[7]), t[8]));
}

// BGK relaxation of fluid cells inside the early window: t -> c
template <bool STRICT, int NP>
__device__ __forceinline__ void vrelax(const PV<NP> (&t)[Q], const PV<NP>& rho, const PV<NP>& mx, const PV<NP>& my, float omega,
                                       PV<NP> (&c)[Q])
{
    PV<NP> ux, uy;
    vdiv2(mx, my, rho, ux, uy);
    const PV<NP> one = vbc<NP>(1.f), om = vbc<NP>(omega);
    PV<NP> d[Q];
    if constexpr (STRICT) {
        // SerialCode/d2q9-bgk.c:349-401 in the reference's operation order (see collide_cell for the identities)
        const PV<NP> uxx = vmul_rounded(ux, ux), uyy = vmul_rounded(uy, uy);
        const PV<NP> u_sq = vadd(uxx, uyy);
        const PV<NP> u5 = vadd(ux, uy), u6 = vsub(uy, ux);
        const PV<NP> v = vdiv_const(u_sq, LBM_2CSQ, LBM_R_2CSQ);
        const PV<NP> q1 = vdiv_const(ux, LBM_C_SQ, LBM_R_C_SQ), q2 = vdiv_const(uy, LBM_C_SQ, LBM_R_C_SQ);
        const PV<NP> q5 = vdiv_const(u5, LBM_C_SQ, LBM_R_C_SQ), q6 = vdiv_const(u6, LBM_C_SQ, LBM_R_C_SQ);
        const PV<NP> s1 = vdiv_const(uxx, LBM_2CSQ2, LBM_R_2CSQ2), s2 = vdiv_const(uyy, LBM_2CSQ2, LBM_R_2CSQ2);
        const PV<NP> s5 = vdiv_const(vmul(u5, u5), LBM_2CSQ2, LBM_R_2CSQ2), s6 = vdiv_const(vmul(u6, u6), LBM_2CSQ2, LBM_R_2CSQ2);
        const PV<NP> w0r = vmul(vbc<NP>(LBM_W0), rho), w1r = vmul(vbc<NP>(LBM_W1), rho), w2r = vmul(vbc<NP>(LBM_W2), rho);
        d[0] = vmul_rounded(w0r, vsub(one, v));
        d[1] = vmul_rounded(w1r, vsub(vadd(vadd(one, q1), s1), v));
        d[3] = vmul_rounded(w1r, vsub(vadd(vsub(one, q1), s1), v));
        d[2] = vmul_rounded(w1r, vsub(vadd(vadd(one, q2), s2), v));
        d[4] = vmul_rounded(w1r, vsub(vadd(vsub(one, q2), s2), v));
        d[5] = vmul_rounded(w2r, vsub(vadd(vadd(one, q5), s5), v));
        d[7] = vmul_rounded(w2r, vsub(vadd(vsub(one, q5), s5), v));
        d[6] = vmul_rounded(w2r, vsub(vadd(vadd(one, q6), s6), v));
        d[8] = vmul_rounded(w2r, vsub(vadd(vsub(one, q6), s6), v));
#pragma unroll
        for (int k = 0; k < Q; k++) c[k] = vadd(t[k], vmul_rounded(om, vsub(d[k], t[k])));
    } else {
        // collide_cell<false>'s formula: fused multiply-adds, multiplications by RN(1/c)
        const PV<NP> u_sq = vfma(ux, ux, vmul(uy, uy));
        const PV<NP> base = vfma(vbc<NP>(-LBM_R_2CSQ), u_sq, one);
        const PV<NP> u5 = vadd(ux, uy), u6 = vsub(uy, ux);
        const PV<NP> k2 = vbc<NP>(LBM_R_2CSQ2), k1 = vbc<NP>(LBM_R_C_SQ), nk1 = vbc<NP>(-LBM_R_C_SQ);
        const PV<NP> e1 = vfma(vmul(k2, ux), ux, base);
        const PV<NP> e2 = vfma(vmul(k2, uy), uy, base);
        const PV<NP> e5 = vfma(vmul(k2, u5), u5, base);
        const PV<NP> e6 = vfma(vmul(k2, u6), u6, base);
        const PV<NP> w0r = vmul(vbc<NP>(LBM_W0), rho), w1r = vmul(vbc<NP>(LBM_W1), rho), w2r = vmul(vbc<NP>(LBM_W2), rho);
        d[0] = vmul(w0r, base);
        d[1] = vmul(w1r, vfma(k1, ux, e1));
        d[3] = vmul(w1r, vfma(nk1, ux, e1));
        d[2] = vmul(w1r, vfma(k1, uy, e2));
        d[4] = vmul(w1r, vfma(nk1, uy, e2));
        d[5] = vmul(w2r, vfma(k1, u5, e5));
        d[7] = vmul(w2r, vfma(nk1, u5, e5));
        d[6] = vmul(w2r, vfma(k1, u6, e6));
        d[8] = vmul(w2r, vfma(nk1, u6, e6));
#pragma unroll
        for (int k = 0; k < Q; k++) c[k] = vfma(om, vsub(d[k], t[k]), t[k]);
    }
}

// the part of the collision after the early window test, for NP pairs side by side: relaxation, bounce-back select,
// |u| of the new state; returns false if a cell left the late window (its |u| has to be redone by the caller)
template <bool STRICT, int NP>
__device__ __forceinline__ bool collide_block(const PV<NP> (&t)[Q], const PV<NP>& rho, const PV<NP>& mx, const PV<NP>& my, uint32_t obits,
                                              float omega, PV<NP> (&o)[Q], float (&speed)[2 * NP])
{
    constexpr int mirror[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    PV<NP> c[Q];
    vrelax<STRICT>(t, rho, mx, my, omega, c);
    // |u| of the stored values
    PV<NP> r2, nx_, ny_;
    vmoments<STRICT>(c, r2, nx_, ny_);
    const bool late = all_in_window(r2, nx_, ny_, obits, 4.76837158203125e-07f /* 2^-21 */, 2097152.f /* 2^21 */, 4.f);
    if constexpr (STRICT) {
        PV<NP> vx, vy;
        vdiv2(nx_, ny_, r2, vx, vy);
        vspeed_from_sq(vadd(vmul_rounded(vx, vx), vmul_rounded(vy, vy)), speed);
    } else {
        vspeed_from_sq(vfma(nx_, nx_, vmul(ny_, ny_)), speed);
#pragma unroll
        for (int h = 0; h < NP; h++) {
            float r0, r1;
            upk(r2.v[h], r0, r1);
            speed[2 * h] = __fdividef(speed[2 * h], r0), speed[2 * h + 1] = __fdividef(speed[2 * h + 1], r1);
        }
    }
    // obstacle: bounce-back permutation (speed 0 keeps the streamed value)
#pragma unroll
    for (int h = 0; h < NP; h++) {
        const bool solid0 = (obits >> (2 * h)) & 1u, solid1 = (obits >> (2 * h + 1)) & 1u;
#pragma unroll
        for (int k = 0; k < Q; k++) {
            float c0, c1, m0, m1;
            upk(c[k].v[h], c0, c1);
            upk(t[mirror[k]].v[h], m0, m1);
            o[k].v[h] = pk(solid0 ? m0 : c0, solid1 ? m1 : c1);
        }
    }
    return late;
}

// Four cells of a thread: t[k][j] = what cell j pulls from plane k (cells 0,1 and 2,3 form the two pairs)  ->  new
// populations o (bounce-back applied to obstacle cells, SerialCode/d2q9-bgk.c:287-299) and |u| of the new state
// (SerialCode:425-452; garbage for obstacle cells, which the callers do not count).
// VERT: both pairs side by side through every operation (more registers: kernels with >= 130 of them), otherwise
// one pair after the other -- still one basic block, so ptxas overlaps the end of the first with the start of the second.
template <bool STRICT, bool VERT>
__device__ __forceinline__ void collide4(const float (&t)[Q][4], uint32_t obits, float omega, float (&o)[Q][4], float (&speed)[4])
{
    obits &= 0xfu;
    PV<2> tp[Q], rho, mx, my;
#pragma unroll
    for (int k = 0; k < Q; k++) tp[k].v[0] = pk(t[k][0], t[k][1]), tp[k].v[1] = pk(t[k][2], t[k][3]);
    vmoments<STRICT>(tp, rho, mx, my);
    // an obstacle cell's collision is discarded: its operands may be anything
    if (!all_in_window(rho, mx, my, obits, 9.5367431640625e-07f /* 2^-20 */, 1048576.f /* 2^20 */, 2.f)) {
        // a fluid cell is outside the window (never in a physical flow): update_cell()'s guarded code, all four
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float tj[Q], oc[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) tj[k] = t[k][j];
            speed[j] = update_cell<STRICT>(tj, (obits >> j) & 1u, omega, oc);
#pragma unroll
            for (int k = 0; k < Q; k++) o[k][j] = oc[k];
        }
        return;
    }
    bool late;
    if constexpr (VERT) {
        PV<2> op[Q];
        late = collide_block<STRICT, 2>(tp, rho, mx, my, obits, omega, op, speed);
#pragma unroll
        for (int k = 0; k < Q; k++) upk(op[k].v[0], o[k][0], o[k][1]), upk(op[k].v[1], o[k][2], o[k][3]);
    } else {
        late = true;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            PV<1> t1[Q], o1[Q], r1, a1, b1;
            float sp[2];
#pragma unroll
            for (int k = 0; k < Q; k++) t1[k].v[0] = tp[k].v[h];
            r1.v[0] = rho.v[h], a1.v[0] = mx.v[h], b1.v[0] = my.v[h];
            late = late & collide_block<STRICT, 1>(t1, r1, a1, b1, (obits >> (2 * h)) & 3u, omega, o1, sp);
            speed[2 * h] = sp[0], speed[2 * h + 1] = sp[1];
#pragma unroll
            for (int k = 0; k < Q; k++) upk(o1[k].v[0], o[k][2 * h], o[k][2 * h + 1]);
        }
    }
    // the new populations are exact; a |u| whose operands left the late window is redone with the guarded code
    if (!late) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float cj[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) cj[k] = o[k][j];
            if (!((obits >> j) & 1u)) {
                if constexpr (STRICT) speed[j] = speed_strict(cj);
                else speed[j] = speed_fast_guarded(cj);
            }
        }
    }
}

} // namespace lbm
