// CPU emulation of the five f-k passes + mask builders: runs the SAME __host__ __device__
// kernel bodies as the GPU (fk_kernels.cuh) block by block with nthr = 1.
// usage: fk_pipeline_emul in.bin out.bin
//   in.bin : int32 nx, ns, kind, taper, col_lo, col_hi ; float64 kval, fval, c0..c3 ;
//            float32 x[nx*ns] ; (kind 1: float64 H[ns]) ; (kind 2: float32 mask[nx*ns])
//   out.bin: float32 y[nx*ns] ; int32 nact
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../das4whales_b200/csrc/fk_hostplan.hpp"
#include "../../das4whales_b200/csrc/fk_kernels.cuh"
using namespace d4w;

template <int T1> static void split_all(bool inv, float2* w, size_t ldw, int t2, const float2* twT, int nact) {
    for (int s = 0; s < nact; ++s)
        for (int t = 0; t < t2; ++t) {
            if (inv) body_row_split<T1, true>(w, ldw, t2, twT, s, t);
            else body_row_split<T1, false>(w, ldw, t2, twT, s, t);
        }
}
template <int X1> static void colA_all(bool inv, const Col2Params& c2p, const float* x, cpd* v2, float* y, const float* taper) {
    for (int c2 = 0; c2 < c2p.x2; ++c2)
        for (int t4 = 0; t4 < c2p.tpn / 2; ++t4) {
            if (inv) body_colA_inv<X1>(c2p, v2, y, c2, t4, c2p.tpb);
            else body_colA_fwd<X1>(c2p, x, v2, taper, c2, t4, c2p.tpb);
        }
}
static void colA_dispatch(bool inv, const Col2Params& c2p, const float* x, cpd* v2, float* y, const float* taper) {
    switch (c2p.x1) {
        case 25: colA_all<25>(inv, c2p, x, v2, y, taper); break;
        case 20: colA_all<20>(inv, c2p, x, v2, y, taper); break;
        case 10: colA_all<10>(inv, c2p, x, v2, y, taper); break;
        default: colA_all<16>(inv, c2p, x, v2, y, taper); break;
    }
}
static void split_dispatch(int t1, bool inv, float2* w, size_t ldw, int t2, const float2* twT, int nact) {
    switch (t1) {
#define C(T) case T: split_all<T>(inv, w, ldw, t2, twT, nact); break;
        C(2) C(3) C(4) C(5) C(6) C(8) C(10) C(12) C(15) C(16) C(20) C(25)
#undef C
        default: fprintf(stderr, "bad t1\n"); exit(2);
    }
}

template <bool INV>
static bool colB_fused_emul(const FkHostPlan& hp, const Col2Params& c2p, cpd* v2, float2* w, size_t ldw, const int2* need, int pl, int tb, cpd* smem) {
#define EM_FUSED3(R0, R1, R2)                                                                                \
    if (hp.fused3 && hp.r3[0] == R0 && hp.r3[1] == R1 && hp.r3[2] == R2) {                                   \
        if constexpr (!INV) body_colB3_fwd<R0, R1, R2>(c2p, v2, w, ldw, need, pl, tb, 0, 1, smem, c2p.tpb, c2p.tpn);           \
        else body_colB3_inv<R0, R1, R2>(c2p, v2, w, ldw, need, pl, tb, 0, 1, smem, c2p.tpb, c2p.tpn);                          \
        return true;                                                                                         \
    }
    EM_FUSED3(10, 10, 10) EM_FUSED3(5, 5, 4)
#undef EM_FUSED3
#define EM_FUSED(RA, RB)                                                                                     \
    if (hp.fused_ra == RA && hp.fused_rb == RB) {                                                            \
        if constexpr (!INV) body_colB_fwd_fused<RA, RB>(c2p, v2, w, ldw, need, pl, tb, 0, 1, smem, c2p.tpb, c2p.tpn);          \
        else body_colB_inv_fused<RA, RB>(c2p, v2, w, ldw, need, pl, tb, 0, 1, smem, c2p.tpb, c2p.tpn);                         \
        return true;                                                                                         \
    }
    EM_FUSED(16, 16) EM_FUSED(16, 20) EM_FUSED(16, 25) EM_FUSED(20, 16) EM_FUSED(20, 20) EM_FUSED(20, 25)
    EM_FUSED(25, 16) EM_FUSED(25, 20) EM_FUSED(25, 25)
#undef EM_FUSED
    return false;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    int hdr[6]; double par[6];
    if (fread(hdr, 4, 6, fi) != 6 || fread(par, 8, 6, fi) != 6) return 3;
    const int nx = hdr[0], ns = hdr[1], kind = hdr[2], taper = hdr[3];
    std::vector<float> x((size_t)nx * ns), dense;
    std::vector<double> H;
    if (fread(x.data(), 4, x.size(), fi) != x.size()) return 3;
    if (kind == MASK_HYBRID_NINF) { H.resize(ns); if (fread(H.data(), 8, ns, fi) != (size_t)ns) return 3; }
    if (kind == MASK_DENSE) { dense.resize((size_t)nx * ns); if (fread(dense.data(), 4, dense.size(), fi) != dense.size()) return 3; }
    fclose(fi);

    FkHostPlan hp; std::string err;
    if (build_fk_hostplan(nx, ns, 227 * 1024, hp, err)) { fprintf(stderr, "plan: %s\n", err.c_str()); return 4; }
    ColParams cp{}; cp.pl = hp.colpl; cp.tw = hp.tw_col.data(); cp.k2pos = hp.k2pos.data(); cp.pos2k = hp.pos2k.data();
    cp.nx = nx; cp.ns = ns; cp.nc = hp.nc; cp.nc_shift = hp.nc_shift; cp.fstride = hp.fstride; cp.aligned = hp.aligned;
    cp.dual = hp.dual; cp.npair = hp.npair; cp.npair_shift = hp.npair_shift; cp.aligned16 = hp.aligned16;
    RowParams rp{}; rp.pl = hp.rowpl; rp.tw = hp.tw_row.data(); rp.twT = hp.twT.data(); rp.t1 = hp.t1; rp.t2 = hp.t2; rp.dual = hp.row_dual;
    MaskParams mp{}; mp.kind = kind; mp.nx = nx; mp.ns = ns; mp.kval = par[0]; mp.fval = par[1];
    mp.c0 = par[2]; mp.c1 = par[3]; mp.c2 = par[4]; mp.c3 = par[5];
    mp.h = H.data(); mp.col_lo = hdr[4]; mp.col_hi = hdr[5]; mp.dense = dense.data();

    // support scan
    const int nrows = nx / 2 + 1;
    std::vector<unsigned int> rowmax(nrows, 0u);
    for (int k = 0; k < nrows; ++k) body_mask_rowmax(mp, rowmax.data(), k, 0, ns, 0, 1);
    std::vector<int> act, k2slot(nrows, -1);
    for (int k = 0; k < nrows; ++k) { float v; memcpy(&v, &rowmax[k], 4); if (v > 0.f) { k2slot[k] = (int)act.size(); act.push_back(k); } }
    const int nact = (int)act.size();
    std::vector<int2> slot_pos(std::max(nact, 1));
    for (int sl = 0; sl < nact; ++sl) slot_pos[sl] = make_int2(hp.k2pos[act[sl]], hp.k2pos[act[sl] == 0 ? 0 : nx - act[sl]]);
    std::vector<float> tab((size_t)std::max(nact, 1) * ns);
    const double scale = 1.0 / ((double)nx * ns);
    for (size_t i = 0; i < (size_t)nact * ns; ++i) body_mask_build(mp, tab.data(), act.data(), hp.pos2k_row_tab.data(), hp.t1, hp.t2, scale, i);

    std::vector<float2> w((size_t)std::max(nact, 1) * ns), smem((size_t)std::max(hp.col_smem, hp.row_smem) / sizeof(float2) + 16);
    std::vector<float> y((size_t)nx * ns, -777.f);
    const size_t ldw = ns;
    const int tile = 2 * hp.nc, ntiles = (ns + tile - 1) / tile;
    const bool two = hp.two_level && nact > 0;
    Col2Params c2p{}; std::vector<int2> need2; std::vector<cpd> v2; std::vector<int> plane_ptr; std::vector<Col2Entry> ents2;
    if (two) {
        c2p.plb = hp.plb; c2p.twb = hp.tw_x2.data(); c2p.twn = hp.tw_col.data(); c2p.nx = nx; c2p.ns = ns; c2p.x1 = hp.x1; c2p.x2 = hp.x2;
        c2p.planes = hp.planes; c2p.np = hp.np2; c2p.fstride = hp.fstride2; c2p.np_shift = hp.np2 == 8 ? 3 : hp.np2 == 4 ? 2 : hp.np2 == 2 ? 1 : 0;
        c2p.vhp = hp.chunk_pairs; c2p.tpb = 0; c2p.tpn = hp.chunk_pairs;
        v2.resize((size_t)hp.planes * hp.x2 * hp.chunk_pairs);
        std::vector<Col2EntryHost> eh; build_col2_entries(hp, k2slot, plane_ptr, eh);
        for (auto& e : eh) ents2.push_back(Col2Entry{e.pos, e.slot, e.flags, 0});
        smem.resize(std::max(smem.size(), hp.colb_smem / sizeof(float2) + 16));
        if (hp.fused_ra || hp.fused3) build_col2_need(hp, k2slot, need2);
        for (int tpb = 0; tpb < ns / 2; tpb += hp.chunk_pairs) {
            c2p.tpb = tpb; c2p.tpn = std::min(hp.chunk_pairs, ns / 2 - tpb);
            colA_dispatch(false, c2p, x.data(), v2.data(), nullptr, taper ? hp.taper.data() : nullptr);
            const int ntb = (c2p.tpn + hp.np2 - 1) / hp.np2;
            for (int pl = 0; pl < hp.planes; ++pl)
                for (int tb = 0; tb < ntb; ++tb)
                    if (!((hp.fused_ra || hp.fused3) && colB_fused_emul<false>(hp, c2p, v2.data(), w.data(), ldw, need2.data(), pl, tb, reinterpret_cast<cpd*>(smem.data()))))
                        body_colB_fwd(c2p, v2.data(), w.data(), ldw, plane_ptr.data(), ents2.data(), pl, tb, 0, 1, reinterpret_cast<cpd*>(smem.data()));
        }
    }
    if (nact && !two)
        for (int b = 0; b < ntiles; ++b) {
            if (hp.dual) body_col_fwd_dual(cp, x.data(), w.data(), ldw, slot_pos.data(), nact, taper ? hp.taper.data() : nullptr, b, 0, 1, reinterpret_cast<cpd*>(smem.data()));
            else body_col_fwd(cp, x.data(), w.data(), ldw, slot_pos.data(), nact, taper ? hp.taper.data() : nullptr, b, 0, 1, smem.data());
        }
    if (hp.t1 > 1 && nact) split_dispatch(hp.t1, false, w.data(), ldw, hp.t2, hp.twT.data(), nact);
    if (hp.row_dual) {
        for (int pr = 0; pr < (nact + 1) / 2; ++pr)
            for (int k1 = 0; k1 < hp.t1; ++k1)
                body_row_mid_dual(rp, w.data(), ldw, tab.data(), (size_t)ns, k1, pr, nact, 0, 1, reinterpret_cast<cpd*>(smem.data()));
    } else {
        for (int s = 0; s < nact; ++s)
            for (int k1 = 0; k1 < hp.t1; ++k1) {
                if (hp.row_fused) body_row_mid_fused(rp, w.data(), ldw, tab.data(), (size_t)ns, k1, s, 0, 1, smem.data());
                else body_row_mid(rp, w.data(), ldw, tab.data(), (size_t)ns, k1, s, 0, 1, smem.data());
            }
    }
    if (hp.t1 > 1 && nact) split_dispatch(hp.t1, true, w.data(), ldw, hp.t2, hp.twT.data(), nact);
    if (two)
        for (int tpb = 0; tpb < ns / 2; tpb += hp.chunk_pairs) {
            c2p.tpb = tpb; c2p.tpn = std::min(hp.chunk_pairs, ns / 2 - tpb);
            const int ntb = (c2p.tpn + hp.np2 - 1) / hp.np2;
            for (int pl = 0; pl < hp.planes; ++pl)
                for (int tb = 0; tb < ntb; ++tb)
                    if (!((hp.fused_ra || hp.fused3) && colB_fused_emul<true>(hp, c2p, v2.data(), w.data(), ldw, need2.data(), pl, tb, reinterpret_cast<cpd*>(smem.data()))))
                        body_colB_inv(c2p, v2.data(), w.data(), ldw, plane_ptr.data(), ents2.data(), pl, tb, 0, 1, reinterpret_cast<cpd*>(smem.data()));
            colA_dispatch(true, c2p, nullptr, v2.data(), y.data(), nullptr);
        }
    for (int b = 0; b < ntiles && !two; ++b) {
        if (hp.dual) body_col_inv_dual(cp, w.data(), ldw, slot_pos.data(), nact, y.data(), b, 0, 1, reinterpret_cast<cpd*>(smem.data()));
        else body_col_inv(cp, w.data(), ldw, slot_pos.data(), nact, y.data(), b, 0, 1, smem.data());
    }

    FILE* fo = fopen(argv[2], "wb");
    fwrite(y.data(), 4, y.size(), fo);
    fwrite(&nact, 4, 1, fo);
    int info[4] = {hp.t1, hp.t2, hp.nc, hp.two_level ? 1000 * (1 + hp.fused_ra) + hp.x1 : (hp.dual ? 100 + hp.colpl.nstages : hp.colpl.nstages)};
    fwrite(info, 4, 4, fo);
    fclose(fo);
    return 0;
}
