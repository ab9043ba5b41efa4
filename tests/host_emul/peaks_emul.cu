// Host emulation of the device peak picker (d4w_find_peaks): the per-sample body peak_pick_one and the block /
// superblock level tables, run on the CPU so that the index logic is covered by the "not gpu" test tier.
//   peaks_emul <in.bin> <out.bin>;  in: int nx, int ns, double thr, float x[nx*ns];  out: uint8 flags[nx*ns]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../das4whales_b200/csrc/rows_kernels.cuh"
using namespace d4w;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    int nx, ns; double thr;
    if (!f || fread(&nx, 4, 1, f) != 1 || fread(&ns, 4, 1, f) != 1 || fread(&thr, 8, 1, f) != 1) return 2;
    std::vector<float> x((size_t)nx * ns);
    if (fread(x.data(), 4, x.size(), f) != x.size()) return 2;
    fclose(f);
    const int nb1 = (ns + kPkB - 1) / kPkB, nb2 = (nb1 + kPkB - 1) / kPkB;
    std::vector<unsigned char> flags((size_t)nx * ns, 0);
    std::vector<float> bm(nb1), bn(nb1), sm(nb2), sn(nb2);
    for (int row = 0; row < nx; ++row) {
        const float* r = x.data() + (size_t)row * ns;
        for (int b = 0; b < nb1; ++b) {                               // what k_peak_levels computes
            float mx = -INFINITY, mn = INFINITY;
            for (int i = b * kPkB; i < ns && i < (b + 1) * kPkB; ++i) { mx = fmaxf(mx, r[i]); mn = fminf(mn, r[i]); }
            bm[b] = mx; bn[b] = mn;
        }
        float rowmin = INFINITY;
        for (int sb = 0; sb < nb2; ++sb) {
            float mx = -INFINITY, mn = INFINITY;
            for (int b = sb * kPkB; b < nb1 && b < (sb + 1) * kPkB; ++b) { mx = fmaxf(mx, bm[b]); mn = fminf(mn, bn[b]); }
            sm[sb] = mx; sn[sb] = mn; rowmin = fminf(rowmin, mn);
        }
        if (ns >= 3)
            for (int i = 0; i < ns; ++i) {
                const int p = peak_pick_one(r, ns, i, bm.data(), bn.data(), sm.data(), sn.data(), nb1, rowmin, thr);
                if (p >= 0) flags[(size_t)row * ns + p] = 1;
            }
    }
    FILE* o = fopen(argv[2], "wb");
    fwrite(flags.data(), 1, flags.size(), o);
    fclose(o);
    return 0;
}
