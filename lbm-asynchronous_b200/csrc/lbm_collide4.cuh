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
// Structure (as before, now per pair): update_cell() guards each of its special sequences (division by rho, the
// constant divisions, the square root) with its own test and branch; ptxas schedules within basic blocks, so here the
// moments of the four cells are formed first and tested against an "early" window; a thread with a cell outside it
// (never in a physical flow) takes update_cell()'s guarded code for its four cells; everybody else runs the sequences
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

// a1 / b and a2 / b on both halves: div2_rn()'s sequence without its operand test (the caller's window implies it)
__device__ __forceinline__ void div2_pair(f2 a1, f2 a2, f2 b, f2& q1, f2& q2)
{
    float b0, b1, y00, y01;
    upk(b, b0, b1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y00) : "f"(b0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y01) : "f"(b1));
    const f2 y0 = pk(y00, y01), nb = neg2(b);
    const f2 e = fma2(nb, y0, bc2(1.f));
    const f2 y = fma2(y0, e, y0);
    const f2 p1 = mul2(a1, y), p2 = mul2(a2, y);
    q1 = fma2(y, fma2(nb, p1, a1), p1);
    q2 = fma2(y, fma2(nb, p2, a2), p2);
}
// div_const() on both halves; nc = -c:  fma(-q, c, x) == fma(q, -c, x) (the product is exact either way)
__device__ __forceinline__ f2 div_const2(f2 x, float c, float rc)
{
    const f2 q = mul2(x, bc2(rc));
    const f2 r = fma2(q, bc2(-c), x);
    return fma2(r, bc2(rc), q);
}
// speed_from_sq()'s sequence on both halves, a select instead of its range test
__device__ __forceinline__ void speed_from_sq_pair(f2 x, float& s0, float& s1)
{
    float x0, x1, y0, y1;
    upk(x, x0, x1);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x0));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(x1));
    const f2 y = pk(y0, y1);
    const f2 g = mul2(x, y), h = mul2(y, bc2(0.5f));
    const f2 s = fma2(fma2(neg2(g), g, x), h, g);
    upk(s, s0, s1);
    s0 = x0 >= 7.888609052210118e-31f /* 2^-100 */ ? s0 : 0.f;
    s1 = x1 >= 7.888609052210118e-31f ? s1 : 0.f;
}
__device__ __forceinline__ bool in_window(float rho, float mx, float my, float lo, float hi, float bound)
{
    return (rho >= lo) & (rho <= hi) & (fmaxf(fabsf(mx), fabsf(my)) <= __fmul_rn(bound, rho)); // no short circuit: no branches
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

// rho and the momentum of a pair of cells; STRICT: SerialCode/d2q9-bgk.c:325-349 (sequential density sum from 0.f,
// velocity brackets left to right)
template <bool STRICT>
__device__ __forceinline__ void moments_pair(const f2 (&t)[Q], f2& rho, f2& mx, f2& my)
{
    f2 d = STRICT ? add2(bc2(0.f), t[0]) : t[0];
#pragma unroll
    for (int k = 1; k < Q; k++) d = add2(d, t[k]);
    rho = d;
    mx = sub2(add2(add2(t[1], t[5]), t[8]), add2(add2(t[3], t[6]), t[7]));
    my = sub2(add2(add2(t[2], t[5]), t[6]), add2(add2(t[4], t[7]), t[8]));
}

// BGK relaxation of a pair of fluid cells inside the early window: t -> c, plus the moments of c
template <bool STRICT>
__device__ __forceinline__ void relax_pair(const f2 (&t)[Q], f2 rho, f2 mx, f2 my, float omega, f2 (&c)[Q])
{
    f2 ux, uy;
    div2_pair(mx, my, rho, ux, uy);
    const f2 one = bc2(1.f), om = bc2(omega);
    f2 d[Q];
    if constexpr (STRICT) {
        // SerialCode/d2q9-bgk.c:349-401 in the reference's operation order (see collide_cell for the identities)
        const f2 uxx = mul2_rounded(ux, ux), uyy = mul2_rounded(uy, uy);
        const f2 u_sq = add2(uxx, uyy);
        const f2 u5 = add2(ux, uy), u6 = sub2(uy, ux);
        const f2 v = div_const2(u_sq, LBM_2CSQ, LBM_R_2CSQ);
        const f2 q1 = div_const2(ux, LBM_C_SQ, LBM_R_C_SQ), q2 = div_const2(uy, LBM_C_SQ, LBM_R_C_SQ);
        const f2 q5 = div_const2(u5, LBM_C_SQ, LBM_R_C_SQ), q6 = div_const2(u6, LBM_C_SQ, LBM_R_C_SQ);
        const f2 s1 = div_const2(uxx, LBM_2CSQ2, LBM_R_2CSQ2), s2 = div_const2(uyy, LBM_2CSQ2, LBM_R_2CSQ2);
        const f2 s5 = div_const2(mul2(u5, u5), LBM_2CSQ2, LBM_R_2CSQ2), s6 = div_const2(mul2(u6, u6), LBM_2CSQ2, LBM_R_2CSQ2);
        const f2 w0r = mul2(bc2(LBM_W0), rho), w1r = mul2(bc2(LBM_W1), rho), w2r = mul2(bc2(LBM_W2), rho);
        d[0] = mul2_rounded(w0r, sub2(one, v));
        d[1] = mul2_rounded(w1r, sub2(add2(add2(one, q1), s1), v));
        d[3] = mul2_rounded(w1r, sub2(add2(sub2(one, q1), s1), v));
        d[2] = mul2_rounded(w1r, sub2(add2(add2(one, q2), s2), v));
        d[4] = mul2_rounded(w1r, sub2(add2(sub2(one, q2), s2), v));
        d[5] = mul2_rounded(w2r, sub2(add2(add2(one, q5), s5), v));
        d[7] = mul2_rounded(w2r, sub2(add2(sub2(one, q5), s5), v));
        d[6] = mul2_rounded(w2r, sub2(add2(add2(one, q6), s6), v));
        d[8] = mul2_rounded(w2r, sub2(add2(sub2(one, q6), s6), v));
#pragma unroll
        for (int k = 0; k < Q; k++) c[k] = add2(t[k], mul2_rounded(om, sub2(d[k], t[k])));
    } else {
        // collide_cell<false>'s formula: fused multiply-adds, multiplications by RN(1/c)
        const f2 u_sq = fma2(ux, ux, mul2(uy, uy));
        const f2 base = fma2(bc2(-LBM_R_2CSQ), u_sq, one);
        const f2 u5 = add2(ux, uy), u6 = sub2(uy, ux);
        const f2 k2 = bc2(LBM_R_2CSQ2), k1 = bc2(LBM_R_C_SQ), nk1 = bc2(-LBM_R_C_SQ);
        const f2 e1 = fma2(mul2(k2, ux), ux, base);
        const f2 e2 = fma2(mul2(k2, uy), uy, base);
        const f2 e5 = fma2(mul2(k2, u5), u5, base);
        const f2 e6 = fma2(mul2(k2, u6), u6, base);
        const f2 w0r = mul2(bc2(LBM_W0), rho), w1r = mul2(bc2(LBM_W1), rho), w2r = mul2(bc2(LBM_W2), rho);
        d[0] = mul2(w0r, base);
        d[1] = mul2(w1r, fma2(k1, ux, e1));
        d[3] = mul2(w1r, fma2(nk1, ux, e1));
        d[2] = mul2(w1r, fma2(k1, uy, e2));
        d[4] = mul2(w1r, fma2(nk1, uy, e2));
        d[5] = mul2(w2r, fma2(k1, u5, e5));
        d[7] = mul2(w2r, fma2(nk1, u5, e5));
        d[6] = mul2(w2r, fma2(k1, u6, e6));
        d[8] = mul2(w2r, fma2(nk1, u6, e6));
#pragma unroll
        for (int k = 0; k < Q; k++) c[k] = fma2(om, sub2(d[k], t[k]), t[k]);
    }
}

// Four cells: t[k][j] = what cell j pulls from plane k  ->  new populations o (bounce-back applied to obstacle cells,
// SerialCode/d2q9-bgk.c:287-299) and |u| of the new state (SerialCode:425-452; garbage for obstacle cells, which
// the callers do not count).
template <bool STRICT>
__device__ __forceinline__ void collide4(const float (&t)[Q][4], uint32_t obits, float omega, float (&o)[Q][4], float (&speed)[4])
{
    f2 tp[2][Q], rho[2], mx[2], my[2];
    bool early = true;
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
        for (int k = 0; k < Q; k++) tp[h][k] = pk(t[k][2 * h], t[k][2 * h + 1]);
        moments_pair<STRICT>(tp[h], rho[h], mx[h], my[h]);
        float r0, r1, a0, a1, b0, b1;
        upk(rho[h], r0, r1), upk(mx[h], a0, a1), upk(my[h], b0, b1);
        // an obstacle cell's collision is discarded: its operands may be anything
        early = early & (in_window(r0, a0, b0, 9.5367431640625e-07f /* 2^-20 */, 1048576.f /* 2^20 */, 2.f) | (((obits >> (2 * h)) & 1u) != 0)) &
                (in_window(r1, a1, b1, 9.5367431640625e-07f, 1048576.f, 2.f) | (((obits >> (2 * h + 1)) & 1u) != 0));
    }
    if (!early) {
        // some fluid cell is outside the window (never in a physical flow): update_cell()'s guarded code, all four
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
    bool late = true;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        f2 c[Q];
        relax_pair<STRICT>(tp[h], rho[h], mx[h], my[h], omega, c);
        // |u| of the stored values
        f2 r2, nx_, ny_;
        moments_pair<STRICT>(c, r2, nx_, ny_);
        float r0, r1, a0, a1, b0, b1;
        upk(r2, r0, r1), upk(nx_, a0, a1), upk(ny_, b0, b1);
        const bool solid0 = (obits >> (2 * h)) & 1u, solid1 = (obits >> (2 * h + 1)) & 1u;
        late = late & (in_window(r0, a0, b0, 4.76837158203125e-07f /* 2^-21 */, 2097152.f /* 2^21 */, 4.f) | solid0) &
               (in_window(r1, a1, b1, 4.76837158203125e-07f, 2097152.f, 4.f) | solid1);
        if constexpr (STRICT) {
            f2 vx, vy;
            div2_pair(nx_, ny_, r2, vx, vy);
            speed_from_sq_pair(add2(mul2_rounded(vx, vx), mul2_rounded(vy, vy)), speed[2 * h], speed[2 * h + 1]);
        } else {
            float s0, s1;
            speed_from_sq_pair(fma2(nx_, nx_, mul2(ny_, ny_)), s0, s1);
            speed[2 * h] = __fdividef(s0, r0), speed[2 * h + 1] = __fdividef(s1, r1);
        }
        // obstacle: bounce-back permutation (speed 0 keeps the streamed value); selected here, per pair, so that the
        // streamed-in values die early instead of living to the end of the block
        constexpr int mirror[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
#pragma unroll
        for (int k = 0; k < Q; k++) {
            float c0, c1;
            upk(c[k], c0, c1);
            o[k][2 * h] = solid0 ? t[mirror[k]][2 * h] : c0;
            o[k][2 * h + 1] = solid1 ? t[mirror[k]][2 * h + 1] : c1;
        }
    }
    // ---- the new populations are exact; a |u| whose operands left the late window is redone with the guarded code
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
