// Truncated Taylor "jets" through the reference's activation functions, forward and reverse.
//
// The reference obtains D_x^k U (k <= n) and D_t^m U by nesting torch.autograd.grad
// (reference Code/PDE_Residual.py:92-97, 109-125, 129-145).  Only *pure* derivatives are ever
// taken (column 1 for x, column 0 for t), never mixed ones, so the same numbers are the
// normalised Taylor coefficients c_k = f^(k)/k! of two univariate series per neuron that
// share their order-0 term: an x-series of order NX and a t-series of order NT.
//
// A jet is a flat array of J = 1 + NX + NT scalars:
//     [0]            order 0 (shared)
//     [1 .. NX]      x-series orders 1..NX
//     [NX+1 .. NX+NT] t-series orders 1..NT
//
// Everything is templated on the scalar type so the very same code is compiled for the device
// (float) and, by tests/jet_harness, for the host (double) where it is checked against
// torch.autograd.  All loop bounds are compile-time constants so that jets live in registers.
#pragma once

#if defined(__CUDACC__)
#define PDR_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define PDR_HD inline
#endif

namespace pdr {

enum Activation : int { ACT_TANH = 0, ACT_SIN = 1, ACT_RATIONAL = 2 };

template <int NX, int NT>
struct JetShape {
    static constexpr int J = 1 + NX + NT;
};

// index of coefficient k of series S (0 = x, 1 = t) inside the flat jet
template <int NX>
PDR_HD constexpr int jet_index(int S, int k) { return k == 0 ? 0 : (S == 0 ? k : NX + k); }

// ---------------------------------------------------------------------------------------------
// scalar helpers (overloaded so the float path uses the single-precision CUDA math functions)
// ---------------------------------------------------------------------------------------------
// float tanh: branch-free, two MUFU ops.  |x| < 0.25: odd Taylor polynomial through x^11 (truncation
// < 1e-10); otherwise 1 - 2/(exp(2|x|) + 1) with ex2.approx / rcp.approx (absolute error ~1.5e-7).
PDR_HD float pdr_tanh(float v) {
#if defined(__CUDA_ARCH__)
    const float ax = fabsf(v);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    const float big = copysignf(fmaf(-2.f, r, 1.f), v);
    const float x2 = v * v;
    float p = fmaf(x2, -0.0088632355f, 0.021869489f);
    p = fmaf(x2, p, -0.053968254f);
    p = fmaf(x2, p, 0.13333333f);
    p = fmaf(x2, p, -0.33333333f);
    const float small = fmaf(v * x2, p, v);
    return ax < 0.25f ? small : big;
#else
    return tanhf(v);
#endif
}
PDR_HD double pdr_tanh(double v) { return tanh(v); }
// float reciprocal without the IEEE-division slow path (a subroutine call in every activation): rcp.approx (1 ulp)
// refined by one Newton step, <= 1 ulp from the correctly rounded quotient for normal arguments.  Used for
// 1/Q(z_0) of the rational activation, Q >= its minimum over the real line (> 0 for the reference's coefficients).
PDR_HD float pdr_rcp(float v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.f), r);
#else
    return 1.f / v;
#endif
}
PDR_HD double pdr_rcp(double v) { return 1.0 / v; }
PDR_HD void pdr_sincos(float v, float* s, float* c)    { *s = sinf(v); *c = cosf(v); }
PDR_HD void pdr_sincos(double v, double* s, double* c) { *s = sin(v); *c = cos(v); }

// ---------------------------------------------------------------------------------------------
// truncated Cauchy product on jets and its adjoint
// ---------------------------------------------------------------------------------------------
template <int NX, int NT, typename T>
PDR_HD void jet_mul(const T* a, const T* b, T* c) {
    c[0] = a[0] * b[0];
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            T s = a[0] * b[off + k] + a[off + k] * b[0];
#pragma unroll
            for (int i = 1; i < k; ++i) s += a[off + i] * b[off + k - i];
            c[off + k] = s;
        }
    }
}

// given cbar = dL/dc for c = a*b, accumulate abar += dL/da and bbar += dL/db
template <int NX, int NT, typename T>
PDR_HD void jet_mul_adj(const T* a, const T* b, const T* cbar, T* abar, T* bbar) {
    abar[0] += cbar[0] * b[0];
    bbar[0] += cbar[0] * a[0];
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            const T g = cbar[off + k];
            // i = 0 and i = k terms touch the shared order-0 slot
            abar[0] += g * b[off + k];
            bbar[off + k] += g * a[0];
            abar[off + k] += g * b[0];
            bbar[0] += g * a[off + k];
#pragma unroll
            for (int i = 1; i < k; ++i) {
                abar[off + i] += g * b[off + k - i];
                bbar[off + k - i] += g * a[off + i];
            }
        }
    }
}

// c = a*a: the symmetric half of the Cauchy product
template <int NX, int NT, typename T>
PDR_HD void jet_sqr(const T* a, T* c) {
    c[0] = a[0] * a[0];
    const T a0x2 = T(2) * a[0];
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            T s = a0x2 * a[off + k];
            T h = T(0);
#pragma unroll
            for (int i = 1; 2 * i < k; ++i) h += a[off + i] * a[off + k - i];
            s += T(2) * h;
            if (k % 2 == 0) s += a[off + k / 2] * a[off + k / 2];
            c[off + k] = s;
        }
    }
}

// given cbar = dL/dc for c = a*a, accumulate abar += dL/da = 2 (cbar correlated with a)
template <int NX, int NT, typename T>
PDR_HD void jet_sqr_adj(const T* a, const T* cbar, T* abar) {
    abar[0] += T(2) * cbar[0] * a[0];
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            const T g = T(2) * cbar[off + k];
            abar[0] += g * a[off + k];
            abar[off + k] += g * a[0];
#pragma unroll
            for (int i = 1; i < k; ++i) abar[off + i] += g * a[off + k - i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tanh:  y' = (1 - y^2) z'.  With w = 1 - y^2:
//     k y_k = sum_{j=1..k} j z_j w_{k-j},   w_k = -sum_{i=0..k} y_i y_{k-i}  (k >= 1)
// (reference activation: torch.nn.Tanh, Network.py:158-160)
// ---------------------------------------------------------------------------------------------
// Everything in the tanh jet depends on z_0 only through y_0 = tanh(z_0): the series part takes y_0 as given (the
// reverse kernels stash y_0 in the slot of z_0 and never evaluate tanh again).  z[0] is not read.
template <int NX, int NT, typename T>
PDR_HD void tanh_series(T y0, const T* z, T* y, T* w) {
    y[0] = y0;
    w[0] = T(1) - y[0] * y[0];
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            T acc = T(0);
#pragma unroll
            for (int j = 1; j <= k; ++j) {
                const T wkj = (k - j == 0) ? w[0] : w[off + k - j];
                acc += (T(j) * z[off + j]) * wkj;
            }
            y[off + k] = acc * (T(1) / T(k));
            if (k < N) {
                T s = T(2) * y[0] * y[off + k];
#pragma unroll
                for (int i = 1; i < k; ++i) s += y[off + i] * y[off + k - i];
                w[off + k] = -s;
            } else {
                w[off + k] = T(0);   // never read
            }
        }
    }
}

template <int NX, int NT, typename T>
PDR_HD void tanh_fwd(const T* z, T* y, T* w) {
    tanh_series<NX, NT, T>(pdr_tanh(z[0]), z, y, w);
}

// w = 1 - y^2 as a jet, from a known y jet
template <int NX, int NT, typename T>
PDR_HD void tanh_w_from_y(const T* y, T* w) {
    w[0] = T(1) - y[0] * y[0];
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            if (k < N) {
                T s = T(2) * y[0] * y[off + k];
#pragma unroll
                for (int i = 1; i < k; ++i) s += y[off + i] * y[off + k - i];
                w[off + k] = -s;
            } else {
                w[off + k] = T(0);   // never read
            }
        }
    }
}

// adjoint of the series recurrences given the forward jets y, w (z[0] is not read)
template <int NX, int NT, typename T>
PDR_HD void tanh_vjp_core(const T* z, const T* y, const T* w, const T* ybar_in, T* zbar) {
    constexpr int J = 1 + NX + NT;
    T yb[J], wb[J];
#pragma unroll
    for (int i = 0; i < J; ++i) { yb[i] = ybar_in[i]; wb[i] = T(0); zbar[i] = T(0); }
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = N; k >= 1; --k) {
            if (k < N) {
                // w_k = -(2 y_0 y_k + sum_{i=1..k-1} y_i y_{k-i})
                const T g = -wb[off + k];
                yb[0] += T(2) * g * y[off + k];
                yb[off + k] += T(2) * g * y[0];
#pragma unroll
                for (int i = 1; i < k; ++i) yb[off + i] += T(2) * g * y[off + k - i];
            }
            // y_k = (1/k) sum_j j z_j w_{k-j}
            const T g = yb[off + k] * (T(1) / T(k));
#pragma unroll
            for (int j = 1; j <= k; ++j) {
                const T wkj = (k - j == 0) ? w[0] : w[off + k - j];
                zbar[off + j] += g * T(j) * wkj;
                if (k - j == 0) wb[0] += g * T(j) * z[off + j];
                else            wb[off + k - j] += g * T(j) * z[off + j];
            }
        }
    }
    // w_0 = 1 - y_0^2 ;  y_0 = tanh z_0
    yb[0] += -T(2) * y[0] * wb[0];
    zbar[0] = yb[0] * w[0];
}

template <int NX, int NT, typename T>
PDR_HD void tanh_vjp(const T* z, const T* ybar_in, T* zbar) {
    constexpr int J = 1 + NX + NT;
    T y[J], w[J];
    tanh_fwd<NX, NT, T>(z, y, w);
    tanh_vjp_core<NX, NT, T>(z, y, w, ybar_in, zbar);
}

// the same with the forward jet y = tanh(z) supplied by the caller (z[0] is not read)
template <int NX, int NT, typename T>
PDR_HD void tanh_vjp_y(const T* z, const T* y, const T* ybar_in, T* zbar) {
    constexpr int J = 1 + NX + NT;
    T w[J];
    tanh_w_from_y<NX, NT, T>(y, w);
    tanh_vjp_core<NX, NT, T>(z, y, w, ybar_in, zbar);
}

// ---------------------------------------------------------------------------------------------
// sin:  s' = c z', c' = -s z'   (reference activation: Sin.forward, Network.py:66-67)
//     k s_k = sum_{j=1..k} j z_j c_{k-j},   k c_k = -sum_{j=1..k} j z_j s_{k-j}
// ---------------------------------------------------------------------------------------------
template <int NX, int NT, typename T>
PDR_HD void sin_fwd(const T* z, T* s, T* c) {
    pdr_sincos(z[0], &s[0], &c[0]);
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            T as = T(0), ac = T(0);
#pragma unroll
            for (int j = 1; j <= k; ++j) {
                const T zj = T(j) * z[off + j];
                const T ckj = (k - j == 0) ? c[0] : c[off + k - j];
                const T skj = (k - j == 0) ? s[0] : s[off + k - j];
                as += zj * ckj;
                ac += zj * skj;
            }
            s[off + k] = as * (T(1) / T(k));
            c[off + k] = -ac * (T(1) / T(k));
        }
    }
}

template <int NX, int NT, typename T>
PDR_HD void sin_vjp(const T* z, const T* ybar_in, T* zbar) {
    constexpr int J = 1 + NX + NT;
    T s[J], c[J], sb[J], cb[J];
    sin_fwd<NX, NT, T>(z, s, c);
#pragma unroll
    for (int i = 0; i < J; ++i) { sb[i] = ybar_in[i]; cb[i] = T(0); zbar[i] = T(0); }
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = N; k >= 1; --k) {
            const T gs = sb[off + k] * (T(1) / T(k));
            const T gc = -cb[off + k] * (T(1) / T(k));
#pragma unroll
            for (int j = 1; j <= k; ++j) {
                const int r = (k - j == 0) ? 0 : off + k - j;
                const T zj = T(j) * z[off + j];
                zbar[off + j] += T(j) * (gs * c[r] + gc * s[r]);
                cb[r] += gs * zj;
                sb[r] += gc * zj;
            }
        }
    }
    zbar[0] = sb[0] * c[0] - cb[0] * s[0];
}

// ---------------------------------------------------------------------------------------------
// rational (3,2):  y = P(z)/Q(z), P = a0 + a1 z + a2 z^2 + a3 z^3, Q = b0 + b1 z + b2 z^2
// (reference Rational.forward, Network.py:53-57; no pole guard there, none here)
//     series division:  y_k = (P_k - sum_{j=1..k} Q_j y_{k-j}) / Q_0
// ab = {a0,a1,a2,a3,b0,b1,b2}
// ---------------------------------------------------------------------------------------------
template <int NX, int NT, typename T>
PDR_HD void rational_parts(const T* z, const T* ab, T* s2, T* s3, T* P, T* Q) {
    constexpr int J = 1 + NX + NT;
    jet_sqr<NX, NT, T>(z, s2);
    jet_mul<NX, NT, T>(s2, z, s3);
#pragma unroll
    for (int i = 0; i < J; ++i) {
        P[i] = ab[1] * z[i] + ab[2] * s2[i] + ab[3] * s3[i];
        Q[i] = ab[5] * z[i] + ab[6] * s2[i];
    }
    P[0] += ab[0];
    Q[0] += ab[4];
}

template <int NX, int NT, typename T>
PDR_HD void jet_div(const T* P, const T* Q, T rq0, T* y) {
    y[0] = P[0] * rq0;
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = 1; k <= N; ++k) {
            T acc = P[off + k] - Q[off + k] * y[0];
#pragma unroll
            for (int j = 1; j < k; ++j) acc -= Q[off + j] * y[off + k - j];
            y[off + k] = acc * rq0;
        }
    }
}

template <int NX, int NT, typename T>
PDR_HD void rational_fwd(const T* z, const T* ab, T* y) {
    constexpr int J = 1 + NX + NT;
    T s2[J], s3[J], P[J], Q[J];
    rational_parts<NX, NT, T>(z, ab, s2, s3, P, Q);
    jet_div<NX, NT, T>(P, Q, pdr_rcp(Q[0]), y);
}

// Adjoint of y = P(z)/Q(z) given the forward quantities s2 = z^2, s3 = z^3, Q, r = 1/Q_0 and y.
// abbar (7 entries) is accumulated into.
template <int NX, int NT, typename T>
PDR_HD void rational_vjp_core(const T* z, const T* s2, const T* s3, const T* Q, T r, const T* y, const T* ab,
                              const T* ybar, T* zbar, T* abbar) {
    constexpr int J = 1 + NX + NT;
    // adjoint of the series division: solve the transposed Toeplitz system for Pbar,
    // then Qbar = -(y correlated with Pbar)
    T Pb[J], Qb[J];
    T p0 = ybar[0];
    T q0 = T(0);
#pragma unroll
    for (int S = 0; S < 2; ++S) {
        const int N = (S == 0) ? NX : NT;
        const int off = (S == 0) ? 0 : NX;
#pragma unroll
        for (int k = N; k >= 1; --k) {
            T acc = ybar[off + k];
#pragma unroll
            for (int j = 1; j <= N - k; ++j) acc -= Q[off + j] * Pb[off + k + j];
            Pb[off + k] = acc * r;
        }
#pragma unroll
        for (int j = 1; j <= N; ++j) {
            p0 -= Q[off + j] * Pb[off + j];
            T acc = Pb[off + j] * y[0];
#pragma unroll
            for (int k = j + 1; k <= N; ++k) acc += Pb[off + k] * y[off + k - j];
            Qb[off + j] = -acc;
            q0 += Pb[off + j] * y[off + j];
        }
    }
    Pb[0] = p0 * r;
    Qb[0] = -(Pb[0] * y[0] + q0);

    // coefficients
    T ga1 = T(0), ga2 = T(0), ga3 = T(0), gb1 = T(0), gb2 = T(0);
    T s2b[J], s3b[J];
#pragma unroll
    for (int i = 0; i < J; ++i) {
        ga1 += Pb[i] * z[i];
        ga2 += Pb[i] * s2[i];
        ga3 += Pb[i] * s3[i];
        gb1 += Qb[i] * z[i];
        gb2 += Qb[i] * s2[i];
        zbar[i] = ab[1] * Pb[i] + ab[5] * Qb[i];
        s2b[i] = ab[2] * Pb[i] + ab[6] * Qb[i];
        s3b[i] = ab[3] * Pb[i];
    }
    abbar[0] += Pb[0];
    abbar[1] += ga1;
    abbar[2] += ga2;
    abbar[3] += ga3;
    abbar[4] += Qb[0];
    abbar[5] += gb1;
    abbar[6] += gb2;

    jet_mul_adj<NX, NT, T>(s2, z, s3b, s2b, zbar);   // s3 = s2 * z
    jet_sqr_adj<NX, NT, T>(z, s2b, zbar);            // s2 = z * z
}

template <int NX, int NT, typename T>
PDR_HD void rational_vjp(const T* z, const T* ab, const T* ybar, T* zbar, T* abbar) {
    constexpr int J = 1 + NX + NT;
    T s2[J], s3[J], P[J], Q[J], y[J];
    rational_parts<NX, NT, T>(z, ab, s2, s3, P, Q);
    const T r = pdr_rcp(Q[0]);
    jet_div<NX, NT, T>(P, Q, r, y);
    rational_vjp_core<NX, NT, T>(z, s2, s3, Q, r, y, ab, ybar, zbar, abbar);
}

// the same with the forward jet y = P/Q supplied by the caller: the numerator and the series division are skipped
template <int NX, int NT, typename T>
PDR_HD void rational_vjp_y(const T* z, const T* y, const T* ab, const T* ybar, T* zbar, T* abbar) {
    constexpr int J = 1 + NX + NT;
    T s2[J], s3[J], Q[J];
    jet_sqr<NX, NT, T>(z, s2);
    jet_mul<NX, NT, T>(s2, z, s3);
#pragma unroll
    for (int i = 0; i < J; ++i) Q[i] = ab[5] * z[i] + ab[6] * s2[i];
    Q[0] += ab[4];
    rational_vjp_core<NX, NT, T>(z, s2, s3, Q, pdr_rcp(Q[0]), y, ab, ybar, zbar, abbar);
}

// ---------------------------------------------------------------------------------------------
// dispatch on the activation kind (compile-time)
// ---------------------------------------------------------------------------------------------
template <int ACT, int NX, int NT, typename T>
PDR_HD void act_fwd(const T* z, const T* ab, T* y) {
    constexpr int J = 1 + NX + NT;
    if (ACT == ACT_TANH) {
        T w[J];
        tanh_fwd<NX, NT, T>(z, y, w);
    } else if (ACT == ACT_SIN) {
        T c[J];
        sin_fwd<NX, NT, T>(z, y, c);
    } else {
        rational_fwd<NX, NT, T>(z, ab, y);
    }
}

template <int ACT, int NX, int NT, typename T>
PDR_HD void act_vjp(const T* z, const T* ab, const T* ybar, T* zbar, T* abbar) {
    if (ACT == ACT_TANH) {
        tanh_vjp<NX, NT, T>(z, ybar, zbar);
    } else if (ACT == ACT_SIN) {
        sin_vjp<NX, NT, T>(z, ybar, zbar);
    } else {
        rational_vjp<NX, NT, T>(z, ab, ybar, zbar, abbar);
    }
}

}  // namespace pdr
