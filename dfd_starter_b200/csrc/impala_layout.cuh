// Flat parameter / buffer layout of the IMPALA policy (policies/impala.py:47-134 in registration order),
// shared by the fp32 / mma.sync kernel (impala_forward.cu) and the tcgen05 kernel (impala_forward_tc.cu).
#pragma once
#include <stdint.h>

namespace {

struct ConvP { int g, be, w, b, bm, bv, cin, cout; };
struct ImpalaP {
    ConvP feat[3];
    ConvP res[2][3][2];
    int fc_g, fc_be, fc_w, fc_b, fc_bm, fc_bv;
    int wih, whh, bih, bhh;
    int pol_g, pol_be, pol_w, pol_b, pol_bm, pol_bv;
    int A;
    int64_t P;
    int seq_w[16], seq_n[16];     // conv weight segments in execution order (L2 prefetch of the next layer)
    int seq_o[16], seq_cin[16];   // running offset of each segment (seq_o[15] = total) and its input channels
};

inline ImpalaP make_impala(int A) {
    ImpalaP L = {};
    int off = 0, boff = 0;
    const int cin_[3] = {3, 16, 32}, cout_[3] = {16, 32, 32};
    auto conv = [&](ConvP& p, int cin, int cout) {
        p.cin = cin; p.cout = cout;
        p.g = off; off += cin;
        p.be = off; off += cin;
        p.bm = boff; boff += cin;
        p.bv = boff; boff += cin;
        boff += 1;   // num_batches_tracked
        p.w = off; off += cout * cin * 9;
        p.b = off; off += cout;
    };
    for (int s = 0; s < 3; ++s) conv(L.feat[s], cin_[s], cout_[s]);
    for (int blk = 0; blk < 2; ++blk)
        for (int s = 0; s < 3; ++s) {
            conv(L.res[blk][s][0], cout_[s], cout_[s]);
            conv(L.res[blk][s][1], cout_[s], cout_[s]);
        }
    L.fc_g = off; off += 2048;
    L.fc_be = off; off += 2048;
    L.fc_bm = boff; boff += 2048;
    L.fc_bv = boff; boff += 2048;
    boff += 1;
    L.fc_w = off; off += 256 * 2048;
    L.fc_b = off; off += 256;
    L.wih = off; off += 1024 * 257;
    L.whh = off; off += 1024 * 256;
    L.bih = off; off += 1024;
    L.bhh = off; off += 1024;
    L.pol_g = off; off += 256;
    L.pol_be = off; off += 256;
    L.pol_bm = boff; boff += 256;
    L.pol_bv = boff; boff += 256;
    boff += 1;
    L.pol_w = off; off += A * 256;
    L.pol_b = off; off += A;
    L.A = A;
    L.P = off;
    int n = 0;
    int run = 0;
    auto seq = [&](const ConvP& p) { L.seq_w[n] = p.w; L.seq_n[n] = p.cout * p.cin * 9; L.seq_o[n] = run; L.seq_cin[n] = p.cin; run += L.seq_n[n]; ++n; };
    for (int s = 0; s < 3; ++s) {
        seq(L.feat[s]);
        for (int blk = 0; blk < 2; ++blk) { seq(L.res[blk][s][0]); seq(L.res[blk][s][1]); }
    }
    L.seq_w[15] = 0; L.seq_n[15] = 0; L.seq_o[15] = run; L.seq_cin[15] = 0;
    return L;
}


}  // namespace
