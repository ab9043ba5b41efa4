// CPU emulation of the shared-memory FFT engine (same __host__ __device__ code the
// kernels run): forward stages vs a float64 DFT, digit-reversal map, inverse round trip.
#include <cstdio>
#include <cstdlib>
#include <complex>
#include <vector>
#include <random>
#include "../../das4whales_b200/csrc/fft_plan.hpp"
using namespace d4w;
typedef std::complex<double> cd;

static std::vector<cd> dft_ref(const std::vector<cd>& x) {   // O(N * sqrt-ish) via naive; fine for N<=12000
    int n = (int)x.size();
    std::vector<cd> tw(n), y(n);
    for (int j = 0; j < n; ++j) tw[j] = std::polar(1.0, -2 * M_PI * j / n);
    for (int k = 0; k < n; ++k) {
        cd acc = 0; long long idx = 0;
        for (int j = 0; j < n; ++j) { acc += x[j] * tw[idx]; idx += k; if (idx >= n) idx -= n; }
        y[k] = acc;
    }
    return y;
}

int main(int argc, char** argv) {
    int maxr = argc > 1 ? atoi(argv[1]) : 25;
    int sizes[] = {1, 2, 6, 30, 120, 175, 240, 38, 45, 625, 1000, 4096, 5000, 5510, 10000, 12000};
    int bad = 0;
    std::mt19937 rng(7);
    std::normal_distribution<float> nd;
    for (int n : sizes) {
        FftPlan pl; std::string err;
        if (!make_plan(n, maxr, pl, err)) { printf("N=%d plan error %s\n", n, err.c_str()); bad++; continue; }
        auto tw = make_twiddles(n);
        auto p2f = make_pos2freq(pl);
        const int nfft = 2, fstride = n + 3;
        std::vector<float2> s((size_t)nfft * fstride), orig;
        for (auto& v : s) v = make_float2(nd(rng), nd(rng));
        orig = s;
        fft_forward_stages(s.data(), pl, tw.data(), nfft, fstride, 0, 1, 0, pl.nstages);
        double maxerr = 0, maxref = 0;
        for (int f = 0; f < nfft; ++f) {
            std::vector<cd> x(n);
            for (int i = 0; i < n; ++i) x[i] = cd(orig[f * fstride + i].x, orig[f * fstride + i].y);
            auto y = dft_ref(x);
            for (int p = 0; p < n; ++p) {
                cd got(s[f * fstride + p].x, s[f * fstride + p].y);
                maxerr = std::max(maxerr, std::abs(got - y[p2f[p]]));
                maxref = std::max(maxref, std::abs(y[p2f[p]]));
            }
        }
        fft_inverse_stages(s.data(), pl, tw.data(), nfft, fstride, 0, 1, 0, pl.nstages);
        double rt = 0;
        for (int f = 0; f < nfft; ++f) for (int i = 0; i < n; ++i) {
            float2 a = s[f * fstride + i], b = orig[f * fstride + i];
            rt = std::max(rt, (double)std::hypot(a.x / n - b.x, a.y / n - b.y));
        }
        printf("N=%6d stages=", n);
        for (int q = 0; q < pl.nstages; ++q) printf("%d%s", pl.radix[q], q + 1 < pl.nstages ? "x" : "");
        printf("  fwd rel err %.2e  roundtrip abs err %.2e\n", maxerr / std::max(maxref, 1e-30), rt);
        if (maxerr / std::max(maxref, 1e-30) > 2e-6 || rt > 3e-5) bad++;
    }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad ? 1 : 0;
}
