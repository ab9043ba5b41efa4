// fft_plan.hpp -- host-side planning for the shared-memory FFT engine: radix schedule,
// twiddle table (double -> fp32, rounded once) and digit-reversal maps.
#pragma once
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include "fft_smem.cuh"

namespace d4w {

inline const std::vector<int>& inreg_radices() {
    static const std::vector<int> r = {25, 20, 16, 15, 12, 10, 8, 6, 5, 4, 3, 2};
    return r;
}

// minimum-stage split of a 2^a 3^b 5^c number into in-register radices <= max_radix
inline bool split_smooth(int m, int max_radix, std::vector<int>& out) {
    std::map<int, std::vector<int>> memo;
    struct Rec {
        std::map<int, std::vector<int>>& memo; int maxr;
        bool go(int m, std::vector<int>& res) {
            if (m == 1) { res.clear(); return true; }
            auto it = memo.find(m);
            if (it != memo.end()) { res = it->second; return !res.empty(); }
            std::vector<int> best;
            for (int r : inreg_radices()) {
                if (r > maxr || m % r) continue;
                std::vector<int> sub;
                if (!go(m / r, sub)) continue;
                sub.insert(sub.begin(), r);
                if (best.empty() || sub.size() < best.size()) best = sub;
            }
            memo[m] = best;
            res = best;
            return !best.empty();
        }
    } rec{memo, max_radix};
    return rec.go(m, out);
}

// Build the stage schedule for length n.  Primes 7..61 become generic stages (first, where
// the leg spacing is long and contiguous); then even radices (descending), then odd ones, so
// the late short-stride stages are odd strides (bank-conflict free for 8-byte words).
inline bool make_plan(int n, int max_radix, FftPlan& pl, std::string& err) {
    if (n < 1) { err = "FFT length must be >= 1"; return false; }
    std::vector<int> generic;
    int m = n;
    for (int p = 7; p <= 61 && m > 1; p += 2) {
        bool prime = true;
        for (int d = 3; d * d <= p; d += 2) if (p % d == 0) prime = false;
        if (!prime) continue;
        while (m % p == 0) { generic.push_back(p); m /= p; }
    }
    int t = m;
    while (t % 2 == 0) t /= 2;
    while (t % 3 == 0) t /= 3;
    while (t % 5 == 0) t /= 5;
    if (t != 1) {
        err = "unsupported FFT length " + std::to_string(n) + ": prime factor > 61 (Bluestein path not built yet)";
        return false;
    }
    std::vector<int> smooth;
    if (m > 1 && !split_smooth(m, max_radix, smooth)) { err = "cannot factor " + std::to_string(n); return false; }
    std::vector<int> ev, od;
    for (int r : smooth) (r % 2 == 0 ? ev : od).push_back(r);
    std::sort(ev.rbegin(), ev.rend());
    std::sort(od.rbegin(), od.rend());
    std::vector<int> order = generic;
    order.insert(order.end(), ev.begin(), ev.end());
    order.insert(order.end(), od.begin(), od.end());
    if ((int)order.size() > kMaxStages) { err = "too many FFT stages"; return false; }
    pl.n = n;
    pl.nstages = (int)order.size();
    int len = n;
    for (int s = 0; s < kMaxStages; ++s) { pl.radix[s] = 1; pl.sub[s] = 1; }
    for (int s = 0; s < pl.nstages; ++s) { pl.radix[s] = order[s]; pl.sub[s] = len; len /= order[s]; }
    return true;
}

// explicit stage order "20,20,25" (tuning knob); returns false if it does not multiply to n
inline bool make_plan_from_string(int n, const char* spec, FftPlan& pl) {
    std::vector<int> order;
    long long prod = 1;
    for (const char* c = spec; *c;) {
        int v = 0;
        while (*c >= '0' && *c <= '9') { v = v * 10 + (*c - '0'); ++c; }
        if (v < 2) return false;
        order.push_back(v); prod *= v;
        if (*c == ',' || *c == 'x') ++c; else if (*c) return false;
    }
    if (prod != n || (int)order.size() > kMaxStages) return false;
    pl.n = n; pl.nstages = (int)order.size();
    int len = n;
    for (int s = 0; s < kMaxStages; ++s) { pl.radix[s] = 1; pl.sub[s] = 1; }
    for (int s = 0; s < pl.nstages; ++s) { pl.radix[s] = order[s]; pl.sub[s] = len; len /= order[s]; }
    return true;
}

inline std::vector<float2> make_twiddles(int n) {
    std::vector<float2> t((size_t)std::max(n, 1));
    const double two_pi = 6.283185307179586476925286766559;
    for (int j = 0; j < n; ++j) {
        // exact octant reduction keeps the table symmetric and accurate to 0.5 ulp
        double a = two_pi * (double)j / (double)n;
        t[j] = make_float2((float)std::cos(a), (float)(-std::sin(a)));
    }
    return t;
}

inline std::vector<int> make_pos2freq(const FftPlan& pl) {
    std::vector<int> v(pl.n);
    for (int p = 0; p < pl.n; ++p) v[p] = pos_to_freq(pl, p);
    return v;
}

}  // namespace d4w
