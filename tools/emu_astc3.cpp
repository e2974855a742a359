// DEVELOPER TOOL: host experiment for the two-phase ASTC search of astc3.cu.
//   phase 1  model-based error estimate of EVERY (partition slot, block mode):
//              e_line + sum_planes [ D(slot,plane,grid) + S(slot,plane,grid)*qvar(level) ] + colour-quantisation term
//            with D = |L (I - P M) t|^2 (what decimation to that grid loses; M = pinv(P), P = bilinear infill)
//   phase 2  exact evaluation (quantise, infill, least-squares end points, decoded error) of the N best
// and reports PSNR as a function of N, so that N and the model constants can be chosen on the CPU.
// Not part of libcfx.so and never a fallback.
#include "../cuttlefish_b200/csrc/astc_core.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace cfx;
using namespace cfx::astc;

namespace {

struct GridMat {
    int nw;
    std::vector<float> M;      // nw x T   pseudo-inverse of the infill
    std::vector<float> R;      // T x T    I - P M
    std::vector<float> kappa;  // T        sum_j (f_ij/16)^2
};

Built g_built;
int g_bw = 0, g_bh = 0;
std::vector<GridMat> g_mats;

void build_mats(const Ctx& c)
{
    const int T = c.tab.texels;
    const GridInfo* gr = reinterpret_cast<const GridInfo*>(c.blob + c.tab.off_grids);
    g_mats.clear();
    for (uint32_t g = 0; g < c.tab.n_grids; ++g) {
        GridMat gm;
        const int nw = gr[g].nw;
        gm.nw = nw;
        std::vector<double> P(T*nw, 0.0);
        gm.kappa.assign(T, 0.0f);
        for (int i = 0; i < T; ++i) {
            const uint2 inf = tab_u32x2(c, c.tab.off_infill + (g*T + i)*8u);
            for (int k = 0; k < 4; ++k) {
                const int j = (inf.x >> (8*k)) & 0xFF, f = (inf.y >> (8*k)) & 0xFF;
                P[i*nw + j] += f/16.0;
            }
            for (int j = 0; j < nw; ++j) gm.kappa[i] += float(P[i*nw + j]*P[i*nw + j]);
        }
        // A = P^T P + ridge; solve A M = P^T
        std::vector<double> A(nw*nw, 0.0), B(nw*T, 0.0);
        for (int a = 0; a < nw; ++a) {
            for (int b = 0; b < nw; ++b) {
                double s = 0;
                for (int i = 0; i < T; ++i) s += P[i*nw + a]*P[i*nw + b];
                A[a*nw + b] = s + (a == b ? 1e-9 : 0.0);
            }
            for (int i = 0; i < T; ++i) B[a*T + i] = P[i*nw + a];
        }
        for (int col = 0; col < nw; ++col) {
            int piv = col;
            for (int r = col + 1; r < nw; ++r) if (fabs(A[r*nw + col]) > fabs(A[piv*nw + col])) piv = r;
            if (piv != col) {
                for (int k = 0; k < nw; ++k) std::swap(A[piv*nw + k], A[col*nw + k]);
                for (int k = 0; k < T; ++k) std::swap(B[piv*T + k], B[col*T + k]);
            }
            const double inv = 1.0/A[col*nw + col];
            for (int k = 0; k < nw; ++k) A[col*nw + k] *= inv;
            for (int k = 0; k < T; ++k) B[col*T + k] *= inv;
            for (int r = 0; r < nw; ++r) {
                if (r == col) continue;
                const double f = A[r*nw + col];
                if (f == 0.0) continue;
                for (int k = 0; k < nw; ++k) A[r*nw + k] -= f*A[col*nw + k];
                for (int k = 0; k < T; ++k) B[r*T + k] -= f*B[col*T + k];
            }
        }
        gm.M.resize(nw*T);
        for (int k = 0; k < nw*T; ++k) gm.M[k] = float(B[k]);
        gm.R.assign(T*T, 0.0f);
        for (int i = 0; i < T; ++i)
            for (int k = 0; k < T; ++k) {
                double s = i == k ? 1.0 : 0.0;
                for (int j = 0; j < nw; ++j) s -= P[i*nw + j]*B[j*T + k];
                gm.R[i*T + k] = float(s);
            }
        g_mats.push_back(gm);
    }
}

constexpr int FX = 8;

struct Cand { float est, exact; uint32_t slot, mode; float tl, td, ts, tc; };
FILE* g_dump = nullptr;

struct Eval {
    float err;
    uint32_t cl;
    int ep[24];
    uint8_t su[2*kMaxTexels];
};

// exact evaluation of (slot, mode) from per-texel ideal weights tt[plane][i]
void exact_eval(const Ctx& c, const int (*v)[4], const Slot& slot, const ModeInfo& m, bool has_alpha, const float* t0,
    const float* t1, Eval& out)
{
    const uint32_t T = c.tab.texels;
    out.err = 3.0e38f;
    const uint32_t pc = slot.pc;
    const int dc = slot.dual_ch;
    const uint32_t planes = dc >= 0 ? 2u : 1u;
    const uint32_t n_ints = pc*(has_alpha ? 8u : 6u);
    const int avail = 128 - int(m.wbits) - (pc == 1 ? 17 : 29) - (dc >= 0 ? 2 : 0);
    if (n_ints > 18 || avail < 0) return;
    const uint32_t cl = tab_u8(c, c.tab.off_clevel + (n_ints >> 1)*128u + uint32_t(avail));
    if (cl == 0xFF) return;
    out.cl = cl;
    const GridMat& gm = g_mats[m.grid];
    const uint32_t L = m.level, nw = m.nw;
    const float nm1 = float(kWqN[L] - 1);
    for (uint32_t pl = 0; pl < planes; ++pl) {
        const float* tt = pl ? t1 : t0;
        for (uint32_t j = 0; j < nw; ++j) {
            float g = 0;
            for (uint32_t i = 0; i < T; ++i) g += gm.M[j*T + i]*tt[i];
            g = fminf(fmaxf(g, 0.0f), 1.0f);
            const int k = min(max(__float2int_rn(g*nm1), 0), int(kWqN[L]) - 1);
            out.su[j*planes + pl] = uint8_t(tab_u8(c, c.tab.off_wq_val + L*32u + uint32_t(k)));
        }
    }
    int w[kMaxTexels][2];
    for (uint32_t i = 0; i < T; ++i) {
        const uint2 inf = tab_u32x2(c, c.tab.off_infill + (uint32_t(m.grid)*T + i)*8u);
        for (uint32_t pl = 0; pl < 2; ++pl) {
            if (pl >= planes) { w[i][pl] = w[i][0]; continue; }
            uint32_t acc = 8;
            for (int q = 0; q < 4; ++q) acc += ((inf.y >> (8*q)) & 0xFFu)*out.su[((inf.x >> (8*q)) & 0xFFu)*planes + pl];
            w[i][pl] = int(acc >> 4);
        }
    }
    for (uint32_t p = 0; p < pc; ++p) {
        double A = 0, B = 0, C = 0, P[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0}, A2 = 0, B2 = 0, C2 = 0, PD = 0, QD = 0;
        for (uint32_t i = 0; i < T; ++i) {
            if (slot.part[i] != p) continue;
            const int ww = w[i][0], iw = 64 - ww;
            A += iw*iw; B += iw*ww; C += ww*ww;
            for (int k = 0; k < 4; ++k) { P[k] += iw*v[i][k]; Q[k] += ww*v[i][k]; }
            if (dc >= 0) {
                const int w2 = w[i][1], i2 = 64 - w2;
                A2 += i2*i2; B2 += i2*w2; C2 += w2*w2; PD += i2*v[i][dc]; QD += w2*v[i][dc];
            }
        }
        for (uint32_t comp = 0; comp < 8; ++comp) {
            const uint32_t ch = comp & 3u, which = comp >> 2;
            float fA = float(A), fB = float(B), fC = float(C), fP = float(P[ch]), fQ = float(Q[ch]);
            if (dc >= 0 && int(ch) == dc) { fA = float(A2); fB = float(B2); fC = float(C2); fP = float(PD); fQ = float(QD); }
            const float det = fA*fC - fB*fB;
            float val;
            if (fabsf(det) < 1e-4f*(fA + fC)*(fA + fC) + 1e-6f) {
                const float4 e = which ? slot.e1[p] : slot.e0[p];
                val = ch == 0 ? e.x : (ch == 1 ? e.y : (ch == 2 ? e.z : e.w));
            } else {
                val = (which ? (fA*fQ - fB*fP) : (fC*fP - fB*fQ))*(64.0f/float(FX))/det;
            }
            int q = 255;
            if (ch < 3 || has_alpha) {
                const int iv = min(max(__float2int_rn(val), 0), 255);
                const uint32_t rank = tab_u8(c, c.tab.off_cq_near + cl*256u + uint32_t(iv));
                q = int(tab_u8(c, c.tab.off_cq_val + cl*256u + rank));
            }
            out.ep[p*8u + comp] = q;
        }
        int* e = out.ep + p*8u;
        if (e[4] + e[5] + e[6] < e[0] + e[1] + e[2])
            for (int k = 0; k < 4; ++k) std::swap(e[k], e[4 + k]);
    }
    double err = 0;
    for (uint32_t i = 0; i < T; ++i) {
        const int* e = out.ep + slot.part[i]*8u;
        for (int k = 0; k < (has_alpha ? 4 : 3); ++k) {
            const int ww = (k == dc) ? w[i][1] : w[i][0];
            const int d = ((e[k]*FX*(64 - ww) + e[4 + k]*FX*ww + 32) >> 6) - v[i][k];
            err += double(d)*d;
        }
    }
    out.err = float(err);
}

// model constants (tunable from the environment)
float envf(const char* name, float def) { const char* s = getenv(name); return s ? float(atof(s)) : def; }

} // namespace

// Per block: estimates + exact errors for every (slot, mode); for each N in ns[] the SSE of the best exact among the
// N best estimates (after `refine` rounds on that winner) is accumulated into sse_out[k].  out (optional) receives the
// blocks encoded with N = ns[pick].
extern "C" int emu3_run(const float* rgba, uint32_t w, uint32_t h, uint32_t bw, uint32_t bh, const uint32_t* ns, uint32_t n_ns,
    double* sse_out, double* rank_hist /* 64 bins: rank (by estimate) of the exact winner */, uint8_t* out, uint32_t pick,
    uint32_t refine, uint32_t quality, int use_all_modes)
{
    if (g_bw != int(bw) || g_bh != int(bh)) {
        g_built = build_tables(bw, bh); g_bw = bw; g_bh = bh;
        Ctx c0; c0.blob = g_built.blob.data(); c0.tab = g_built.tab;
        build_mats(c0);
    }
    Ctx c; c.blob = g_built.blob.data(); c.tab = g_built.tab;
    const uint32_t T = bw*bh;
    const uint32_t bxn = (w + bw - 1)/bw, byn = (h + bh - 1)/bh;
    const float kD = envf("K_D", 1.0f), kS = envf("K_S", 1.0f), kC = envf("K_C", 1.0f), kLine = envf("K_LINE", 1.0f),
        kRefit = envf("K_REFIT", 0.75f), kQmode = envf("K_QMODE", 1.0f);
    static BlockState st;
    std::vector<Cand> cands;
    if (getenv("EMU3_DUMP")) g_dump = fopen(getenv("EMU3_DUMP"), "w");
    for (uint32_t k = 0; k < n_ns; ++k) sse_out[k] = 0;
    for (uint32_t by = 0; by < byn; ++by)
        for (uint32_t bx = 0; bx < bxn; ++bx) {
            bool constant = true, has_alpha = false;
            int v[kMaxTexels][4];
            for (uint32_t i = 0; i < T; ++i) {
                uint32_t x = std::min(bx*bw + i % bw, w - 1), y = std::min(by*bh + i/bw, h - 1);
                const float* p = rgba + (size_t(y)*w + x)*4;
                auto cl = [](float f) { return std::min(std::max(f, 0.0f), 1.0f)*255.0f; };
                st.cf[i] = make_float4(cl(p[0]), cl(p[1]), cl(p[2]), cl(p[3]));
                v[i][0] = __float2int_rn(st.cf[i].x*FX); v[i][1] = __float2int_rn(st.cf[i].y*FX);
                v[i][2] = __float2int_rn(st.cf[i].z*FX); v[i][3] = __float2int_rn(st.cf[i].w*FX);
                if (st.cf[i].w != 255.0f) has_alpha = true;
                if (memcmp(&st.cf[i], &st.cf[0], 16) != 0) constant = false;
            }
            st.has_alpha = has_alpha;
            const size_t bi = size_t(by)*bxn + bx;
            if (constant) {
                if (out) { uint4 blk = pack_void_extent(st.cf[0]); memcpy(out + bi*16, &blk, 16); }
                continue;
            }
            for (uint32_t i = 0; i < kSlots; ++i) st.slots[i].valid = 0;
            for (uint32_t lane = 0; lane < 32; ++lane) step_init(c, st, lane);
            for (uint32_t lane = 0; lane < 32; ++lane) step_rank(c, st, lane);
            for (uint32_t lane = 0; lane < 32; ++lane) step_score(c, st, lane);
            for (uint32_t lane = 0; lane < 32; ++lane) step_slots(c, st, lane);

            cands.clear();
            const uint32_t nch = has_alpha ? 4u : 3u;
            for (uint32_t s = 0; s < kSlots; ++s) {
                const Slot& slot = st.slots[s];
                if (!slot.valid) continue;
                const uint32_t type = slot_type(s);
                const int dc = slot.dual_ch;
                const uint32_t planes = dc >= 0 ? 2u : 1u;
                // per-texel squared line length of each plane
                float len2[2][kMaxTexels];
                float lo = 1e30f, hi = -1e30f;
                if (dc >= 0) for (uint32_t i = 0; i < T; ++i) { const float xc = ch(st.cf[i], dc); lo = fminf(lo, xc); hi = fmaxf(hi, xc); }
                for (uint32_t i = 0; i < T; ++i) { len2[0][i] = slot.len2[slot.part[i]]; len2[1][i] = (hi - lo)*(hi - lo); }
                // data-dependent quantisation error of the undecimated ideal weights, per level
                float qe[kWeightLevels], sumlen2 = 0.0f;
                for (uint32_t L = 0; L < kWeightLevels; ++L) {
                    qe[L] = 0.0f;
                    const float nm1 = float(kWqN[L] - 1);
                    for (uint32_t pl = 0; pl < planes; ++pl) {
                        const float* tt = pl ? slot.t2 : slot.t;
                        for (uint32_t i = 0; i < T; ++i) {
                            const int k = min(max(__float2int_rn(tt[i]*nm1), 0), int(kWqN[L]) - 1);
                            const float q = float(tab_u8(c, c.tab.off_wq_val + L*32u + uint32_t(k)))*(1.0f/64.0f);
                            qe[L] += len2[pl][i]*(tt[i] - q)*(tt[i] - q);
                        }
                    }
                }
                for (uint32_t pl = 0; pl < planes; ++pl) for (uint32_t i = 0; i < T; ++i) sumlen2 += len2[pl][i];
                // D and S per grid
                std::vector<float> D(c.tab.n_grids, 0.0f), S(c.tab.n_grids, 0.0f);
                for (uint32_t g = 0; g < c.tab.n_grids; ++g) {
                    const GridMat& gm = g_mats[g];
                    for (uint32_t pl = 0; pl < planes; ++pl) {
                        const float* tt = pl ? slot.t2 : slot.t;
                        for (uint32_t i = 0; i < T; ++i) {
                            float r = 0;
                            for (uint32_t k = 0; k < T; ++k) r += gm.R[i*T + k]*tt[k];
                            D[g] += len2[pl][i]*r*r;
                            S[g] += len2[pl][i]*gm.kappa[i];
                        }
                    }
                }
                const uint32_t first = type == 3 ? c.tab.n_modes1 : 0, count = type == 3 ? c.tab.n_modes2 : c.tab.n_modes1;
                const uint32_t nlist = use_all_modes ? count : c.tab.n_cand_q[quality][type];
                for (uint32_t ci = 0; ci < nlist; ++ci) {
                    const uint32_t mi = use_all_modes ? first + ci : tab_u16(c, c.tab.off_cand_q[quality][type] + ci*2u);
                    const ModeInfo m = tab_mode(c, mi);
                    const uint32_t n_ints = slot.pc*(has_alpha ? 8u : 6u);
                    const int avail = 128 - int(m.wbits) - (slot.pc == 1 ? 17 : 29) - (dc >= 0 ? 2 : 0);
                    if (n_ints > 18 || avail < 0) continue;
                    const uint32_t cl = tab_u8(c, c.tab.off_clevel + (n_ints >> 1)*128u + uint32_t(avail));
                    if (cl == 0xFF) continue;
                    const float wstep = 1.0f/float(kWqN[m.level] - 1)*sqrtf(1.0f - kRefit/float(kWqN[m.level] - 1));
                    const float cstep = 255.0f/float(kColorQuant[cl].n - 1);
                    Cand cd;
                    cd.slot = s; cd.mode = mi;
                    const float refit = 1.0f - kRefit/float(kWqN[m.level] - 1);
                    const float qterm = kQmode > 0.5f ? (sumlen2 > 0 ? S[m.grid]/sumlen2*qe[m.level]*refit : 0.0f)
                                                      : S[m.grid]*wstep*wstep*(1.0f/12.0f);
                    cd.est = kLine*slot.e_line + kD*D[m.grid] + kS*qterm +
                        kC*float(T*nch)*cstep*cstep*(1.0f/18.0f);
                    cd.tl = slot.e_line; cd.td = D[m.grid]; cd.ts = qterm;
                    cd.tc = float(T*nch)*cstep*cstep*(1.0f/18.0f);
                    Eval ev;
                    exact_eval(c, v, slot, m, has_alpha, slot.t, slot.t2, ev);
                    cd.exact = ev.err/float(FX*FX);
                    if (g_dump) fprintf(g_dump, "%zu %u %u %u %u %u %g %g %g %g %g\n", bi, s, uint32_t(m.nw), uint32_t(m.level), cl, slot.pc,
                        cd.tl, cd.td, cd.ts, cd.tc, cd.exact);
                    cands.push_back(cd);
                }
            }
            std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.est < b.est; });
            // rank of the exact winner
            size_t win = 0;
            for (size_t k = 1; k < cands.size(); ++k) if (cands[k].exact < cands[win].exact) win = k;
            if (rank_hist) rank_hist[std::min<size_t>(win, 63)] += 1;
            for (uint32_t k = 0; k < n_ns; ++k) {
                const size_t n = std::min<size_t>(ns[k], cands.size());
                size_t b = 0;
                for (size_t q = 1; q < n; ++q) if (cands[q].exact < cands[b].exact) b = q;
                // refine the winner
                const Slot& slot = st.slots[cands[b].slot];
                const ModeInfo m = tab_mode(c, cands[b].mode);
                Eval best;
                exact_eval(c, v, slot, m, has_alpha, slot.t, slot.t2, best);
                for (uint32_t r = 0; r < refine && best.err > 0; ++r) {
                    float tp[2][kMaxTexels];
                    const int dc = slot.dual_ch;
                    for (uint32_t i = 0; i < T; ++i) {
                        const int* e = best.ep + slot.part[i]*8u;
                        const float xs[4] = {st.cf[i].x, st.cf[i].y, st.cf[i].z, st.cf[i].w};
                        float num0 = 0, den0 = 0, num1 = 0, den1 = 0;
                        for (int chn = 0; chn < 4; ++chn) {
                            if (chn == 3 && !has_alpha) continue;
                            const float a = float(e[chn]), d = float(e[4 + chn]) - a;
                            if (chn == dc) { num1 += (xs[chn] - a)*d; den1 += d*d; } else { num0 += (xs[chn] - a)*d; den0 += d*d; }
                        }
                        tp[0][i] = den0 > 0 ? fminf(fmaxf(num0/den0, 0.0f), 1.0f) : 0.0f;
                        tp[1][i] = den1 > 0 ? fminf(fmaxf(num1/den1, 0.0f), 1.0f) : 0.0f;
                    }
                    Eval trial;
                    exact_eval(c, v, slot, m, has_alpha, tp[0], tp[1], trial);
                    if (trial.err < best.err) best = trial; else break;
                }
                sse_out[k] += best.err/float(FX*FX);
                if (out && k == pick) {
                    Enc enc;
                    enc.clevel = best.cl; enc.err = best.err;
                    for (uint32_t s = 0; s < slot.pc; ++s) {
                        const int* e = best.ep + s*8u;
                        enc.ep[s][0] = uint32_t(e[0]) | (uint32_t(e[1]) << 8) | (uint32_t(e[2]) << 16) | (uint32_t(e[3]) << 24);
                        enc.ep[s][1] = uint32_t(e[4]) | (uint32_t(e[5]) << 8) | (uint32_t(e[6]) << 16) | (uint32_t(e[7]) << 24);
                    }
                    uint4 blk = pack_block(c, slot, m, enc, has_alpha, best.su, 0, true);
                    memcpy(out + bi*16, &blk, 16);
                }
            }
        }
    if (g_dump) { fclose(g_dump); g_dump = nullptr; }
    return 0;
}
