// CPU emulation of the prime-factor 2520-point transform (csrc/fft_pfa.cuh): the same __host__ __device__ stage code and
// dual-lane butterflies (radix 5, 7, 8, 9) the matched-filter kernel runs, checked against a float64 DFT:
//   forward spectrum at position p == DFT[pos2freq[p]];  forward -> x 1 -> inverse == 2520 * input (round trip);
//   fused last stage with a table of ones == plain forward + inverse of the last dimension.
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include <random>
#include "../../das4whales_b200/csrc/fft_pfa.cuh"
using namespace d4w;
typedef std::complex<double> cd;

int main() {
    const int n = kPfaN;
    std::vector<int> t2p, p2f;
    pfa_build_maps(t2p, p2f);
    std::mt19937 rng(11);
    std::normal_distribution<float> nd;
    std::vector<cd> xa(n), xb(n);
    std::vector<cpd> S(n), B(n);
    for (int i = 0; i < n; ++i) {
        const float ar = nd(rng), ai = nd(rng), br = nd(rng), bi = nd(rng);
        xa[i] = cd(ar, ai); xb[i] = cd(br, bi);
        S[t2p[i]] = dmake(f2x_set(ar, br), f2x_set(ai, bi));
    }
    // bijectivity of the maps
    std::vector<int> seen(n, 0), seenf(n, 0);
    for (int i = 0; i < n; ++i) { seen[t2p[i]]++; seenf[p2f[i]]++; }
    int bad = 0;
    for (int i = 0; i < n; ++i) if (seen[i] != 1 || seenf[i] != 1) bad++;
    // reference DFT of both lanes
    std::vector<cd> tw(n), Xa(n), Xb(n);
    for (int j = 0; j < n; ++j) tw[j] = std::polar(1.0, -2 * M_PI * j / n);
    for (int k = 0; k < n; ++k) {
        cd a = 0, b = 0; long long idx = 0;
        for (int j = 0; j < n; ++j) { a += xa[j] * tw[idx]; b += xb[j] * tw[idx]; idx += k; if (idx >= n) idx -= n; }
        Xa[k] = a; Xb[k] = b;
    }
    std::vector<cpd> F = S;
    pfa_forward_3(F.data(), 0, 1);
    pfa_stage<9, 1, false>(F.data(), 0, 1);
    double e = 0, ref = 0;
    for (int p = 0; p < n; ++p) {
        const cd ga(f2x_lo(F[p].x), f2x_lo(F[p].y)), gb(f2x_hi(F[p].x), f2x_hi(F[p].y));
        e = std::max(e, std::max(std::abs(ga - Xa[p2f[p]]), std::abs(gb - Xb[p2f[p]])));
        ref = std::max(ref, std::max(std::abs(Xa[p2f[p]]), std::abs(Xb[p2f[p]])));
    }
    printf("forward rel err %.2e\n", e / ref);
    if (e / ref > 2e-6) bad++;
    // the kernel's path: forward over 3 dims, fused last stage with an all-ones table, inverse over 3 dims
    std::vector<float2> ones(n, make_float2(1.f, 0.f));
    std::vector<cpd> G = S;
    pfa_forward_3(G.data(), 0, 1);
    pfa_last_fused(G.data(), B.data(), ones.data(), 0, 1);
    pfa_inverse_3(B.data(), 0, 1);
    double rt = 0;
    for (int i = 0; i < n; ++i) {
        const cpd z = B[t2p[i]];
        const cd ga(f2x_lo(z.x) / n, f2x_lo(z.y) / n), gb(f2x_hi(z.x) / n, f2x_hi(z.y) / n);
        rt = std::max(rt, std::max(std::abs(ga - xa[i]), std::abs(gb - xb[i])));
    }
    printf("round trip abs err %.2e\n", rt);
    if (rt > 2e-5) bad++;
    // table semantics: tab[m * 280 + j] multiplies position j * 9 + m -> a circular shift by 3 samples via its spectrum
    std::vector<float2> tab(n);
    for (int j = 0; j < n / 9; ++j) for (int m = 0; m < 9; ++m) {
        const int k = p2f[j * 9 + m];
        const double a = -2 * M_PI * (double)((3LL * k) % n) / n;
        tab[m * (n / 9) + j] = make_float2((float)cos(a), (float)sin(a));
    }
    G = S;
    pfa_forward_3(G.data(), 0, 1);
    pfa_last_fused(G.data(), B.data(), tab.data(), 0, 1);
    pfa_inverse_3(B.data(), 0, 1);
    double sh = 0;
    for (int i = 0; i < n; ++i) {
        const cpd z = B[t2p[i]];
        const cd ga(f2x_lo(z.x) / n, f2x_lo(z.y) / n);
        sh = std::max(sh, std::abs(ga - xa[(i - 3 + n) % n]));
    }
    printf("shift-by-table abs err %.2e\n", sh);
    if (sh > 3e-5) bad++;
    // first stage fused with the scatter: time-ordered input -> same spectrum
    {
        std::vector<cpd> T(n), S2(n);
        for (int i = 0; i < n; ++i) T[i] = S[t2p[i]];
        pfa_first_from_time(T.data(), S2.data(), 0, 1);
        pfa_forward_23(S2.data(), 0, 1);
        std::vector<cpd> F2 = S;
        pfa_forward_3(F2.data(), 0, 1);
        double df = 0;
        for (int p = 0; p < n; ++p)
            df = std::max(df, (double)std::max(std::fabs(f2x_lo(S2[p].x) - f2x_lo(F2[p].x)), std::fabs(f2x_hi(S2[p].y) - f2x_hi(F2[p].y))));
        printf("fused first stage vs scatter + stage: max abs diff %.2e\n", df);
        if (df > 1e-6) bad++;
    }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad ? 1 : 0;
}
