// fft_pfa.cuh -- prime-factor (Good-Thomas) 2520-point transform on dual-lane (cpd) elements for the overlap-save
// matched filter (detect.compute_cross_correlogram, /root/reference/src/das4whales/detect.py:140-166).
//
// 2520 = 5 * 7 * 8 * 9 with pairwise coprime factors, so the length-2520 DFT is EXACTLY a 4-D DFT of a [5][7][8][9] array:
// no twiddle factors between the stages (in the Cooley-Tukey blocks they were ~45 % of the instructions), every stage a
// batch of small in-register butterflies.  Index maps (N_i = 5, 7, 8, 9; M_i = 2520 / N_i = 504, 360, 315, 280):
//   time      n = (sum_i n_i M_i) mod 2520                    (Ruritanian map)   <=>  n_i = n * (M_i^-1 mod N_i) mod N_i
//   frequency k = (sum_i k_i M_i (M_i^-1 mod N_i)) mod 2520   (CRT map)          <=>  k_i = k mod N_i
// so that W_2520^{n k} = prod_i W_{N_i}^{n_i k_i}.  Element (a, b, c, d) lives at position ((a*7 + b)*8 + c)*9 + d; the
// transform is in place (frequency digit replaces time digit dimension by dimension), forward and inverse share the maps.
#pragma once
#include <vector>
#include "fft_dual.cuh"

namespace d4w {

constexpr int kPfaN = 2520;
constexpr int kPfaDims[4] = {5, 7, 8, 9};
constexpr int kPfaStride[4] = {504, 72, 9, 1};

// host: position of time index n, and frequency held at position p after the forward transform
inline void pfa_build_maps(std::vector<int>& time2pos, std::vector<int>& pos2freq) {
    int inv[4], M[4];
    for (int i = 0; i < 4; ++i) {
        M[i] = kPfaN / kPfaDims[i];
        inv[i] = 1;
        while ((M[i] % kPfaDims[i]) * inv[i] % kPfaDims[i] != 1) ++inv[i];
    }
    time2pos.assign(kPfaN, 0);
    pos2freq.assign(kPfaN, 0);
    for (int n = 0; n < kPfaN; ++n) {
        int p = 0;
        for (int i = 0; i < 4; ++i) p += ((n * inv[i]) % kPfaDims[i]) * kPfaStride[i];
        time2pos[n] = p;
    }
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 7; ++b)
            for (int c = 0; c < 8; ++c)
                for (int d = 0; d < 9; ++d) {
                    const int dig[4] = {a, b, c, d};
                    long long k = 0;
                    for (int i = 0; i < 4; ++i) k += (long long)dig[i] * M[i] * inv[i];
                    pos2freq[((a * 7 + b) * 8 + c) * 9 + d] = (int)(k % kPfaN);
                }
}

// one twiddle-free stage: R-point DFTs along the dimension with stride ST for every combination of the other digits
template <int R, int ST, bool INV>
__host__ __device__ inline void pfa_stage(cpd* __restrict__ s, int tid, int nthr) {
    constexpr int items = kPfaN / R;
    for (int j = tid; j < items; j += nthr) {
        const int hi = j / ST, lo = j - hi * ST;
        cpd* base = s + hi * (ST * R) + lo;
        cpd v[R];
        static_for<R>([&](auto qc) { constexpr int q = decltype(qc)::value; v[q] = base[q * ST]; });
        DFTD<R, INV>::run(v);
        static_for<R>([&](auto qc) { constexpr int q = decltype(qc)::value; base[q * ST] = v[outpos<R>(q)]; });
    }
}

// first forward stage (dimension a, radix 5) fused with the time-order -> position scatter: item (b, c, d) gathers its five
// inputs straight from the time-ordered block T at n_a = (360 b + 315 c + 280 d + 504 a) mod 2520 and writes positions
// (a, b, c, d) of S.  Threads run over c fastest: the gather stride is then 315 elements = 1260 words = 12 (mod 32) and the
// store stride 9 elements = 4 (mod 32) -- both conflict-free for 16-byte accesses of a quarter warp.
__host__ __device__ inline void pfa_first_from_time(const cpd* __restrict__ T, cpd* __restrict__ S, int tid, int nthr) {
    for (int j = tid; j < 504; j += nthr) {
        const int c = j & 7, bd = j >> 3;            // bd = b * 9 + d, 63 combinations
        const int b = bd / 9, d = bd - 9 * b;
        int n = (360 * b + 315 * c + 280 * d) % kPfaN;
        cpd v[5];
        static_for<5>([&](auto ac) {
            constexpr int a = decltype(ac)::value;
            v[a] = T[n];
            n += 504; if (n >= kPfaN) n -= kPfaN;
        });
        DFTD<5, false>::run(v);
        cpd* base = S + b * 72 + c * 9 + d;
        static_for<5>([&](auto qc) { constexpr int q = decltype(qc)::value; base[q * 504] = v[outpos<5>(q)]; });
    }
}

// forward over the first three dimensions (5, 7, 8); the last one (9, contiguous) is fused with the spectrum multiply
__host__ __device__ inline void pfa_forward_23(cpd* s, int tid, int nthr) {       // after pfa_first_from_time
    pfa_stage<7, 72, false>(s, tid, nthr); D4W_SYNC();
    pfa_stage<8, 9, false>(s, tid, nthr); D4W_SYNC();
}
__host__ __device__ inline void pfa_forward_3(cpd* s, int tid, int nthr) {
    pfa_stage<5, 504, false>(s, tid, nthr); D4W_SYNC();
    pfa_stage<7, 72, false>(s, tid, nthr); D4W_SYNC();
    pfa_stage<8, 9, false>(s, tid, nthr); D4W_SYNC();
}
__host__ __device__ inline void pfa_inverse_3(cpd* s, int tid, int nthr) {
    pfa_stage<8, 9, true>(s, tid, nthr); D4W_SYNC();
    pfa_stage<7, 72, true>(s, tid, nthr); D4W_SYNC();
    pfa_stage<5, 504, true>(s, tid, nthr); D4W_SYNC();
}

// last dimension: forward radix 9 -> x table (lane-independent complex scalars, tab[m * 280 + j] for position j * 9 + m)
// -> inverse radix 9, from S to B (S stays intact for the next template)
__host__ __device__ inline void pfa_last_fused(const cpd* __restrict__ S, cpd* __restrict__ B, const float2* __restrict__ tab, int tid, int nthr) {
    constexpr int R = 9, G = kPfaN / R;
    for (int j = tid; j < G; j += nthr) {
        cpd v[R], u[R];
        float2 tb[R];                                   // table loads first: their L2 latency hides under the forward butterfly
        static_for<R>([&](auto mc) { constexpr int m = decltype(mc)::value; tb[m] = tab[m * G + j]; });
        static_for<R>([&](auto qc) { constexpr int q = decltype(qc)::value; v[q] = S[j * R + q]; });
        DFTD<R, false>::run(v);
        static_for<R>([&](auto mc) { constexpr int m = decltype(mc)::value; u[m] = dmul_s(v[outpos<R>(m)], tb[m]); });
        DFTD<R, true>::run(u);
        static_for<R>([&](auto qc) { constexpr int q = decltype(qc)::value; B[j * R + q] = u[outpos<R>(q)]; });
    }
}

}  // namespace d4w
