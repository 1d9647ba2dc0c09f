// Stand-alone timing + cross-check harness for the GRU recurrence kernels of learner.cuh (not product code).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/gru_bench tools/gru_bench.cu
//   tools/gru_bench [R] [TT] [rows checked] [variant 7|8]
// Checks k_gru_fwd7/8 / k_gru_bwd7/8 against an fp64 CPU recurrence on the first rows, then times them.
#include "../ma_league_b200/csrc/gru_rec.cuh"
#include <vector>
#include <random>
#include <cstdlib>
void mal_set_error(const char *, ...) {}

static float *dev(const std::vector<float> &v) {
    float *p; cudaMalloc(&p, v.size() * 4); cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice); return p;
}
template <typename F> static float time_us(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1000.f / reps;
}

int main(int argc, char **argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 160, TT = argc > 2 ? atoi(argv[2]) : 201;
    const int d_in = 64, A = 11; const int NCHK = argc > 3 ? atoi(argv[3]) : 3;
    const AgentLayout L = agent_layout(d_in, A);
    std::mt19937 rng(1);
    std::uniform_real_distribution<float> U(-0.125f, 0.125f);
    std::normal_distribution<float> Nrm(0.f, 1.f);
    std::vector<float> P0(L.total), P1(L.total);
    for (auto &x : P0) x = U(rng);
    for (auto &x : P1) x = U(rng);
    const size_t M = (size_t)TT * R;
    std::vector<float> gi0(M * G3), gi1(M * G3), dhh(M * HID);
    for (auto &x : gi0) x = Nrm(rng);
    for (auto &x : gi1) x = Nrm(rng);
    for (auto &x : dhh) x = 0.01f * Nrm(rng);
    float *dP0 = dev(P0), *dP1 = dev(P1), *dgi0 = dev(gi0), *dgi1 = dev(gi1), *ddhh = dev(dhh);
    float *h[2], *gates, *dg;
    for (int n = 0; n < 2; ++n) { cudaMalloc(&h[n], M * HID * 4); cudaMemset(h[n], 0, M * HID * 4); }
    cudaMalloc(&gates, M * 256 * 4); cudaMalloc(&dg, M * 256 * 4);
    GruFwdArgs fa; fa.params[0] = dP0; fa.params[1] = dP1; fa.gi[0] = dgi0; fa.gi[1] = dgi1; fa.hout[0] = h[0]; fa.hout[1] = h[1];
    fa.gates = gates; fa.TT = TT; fa.R = R; fa.d_in = d_in; fa.n_actions = A; fa.t0 = 0; fa.t1 = TT;
    GruBwdArgs ba; ba.params = dP0; ba.hout = h[0]; ba.gates = gates; ba.dh_head = ddhh; ba.d_g = dg; ba.TT = TT; ba.R = R;
    ba.d_in = d_in; ba.n_actions = A;

    const int variant = argc > 4 ? atoi(argv[4]) : 8;
    // variant 12 = k_gru_fwd9 in balanced mode: 2 * 148 workers of equal length
    int *flags; cudaMalloc(&flags, 4096); cudaMemset(flags, 0, 4096);
    GruFwdArgs fb = fa; fb.bal_chains = 2 * R; fb.bal_D = (int)(((long long)2 * R * TT + 295) / 296); fb.chain_flags = flags;
    auto fwd = [&] { if (variant == 7) k_gru_fwd7<0><<<dim3(R, 2), HID>>>(fa); else if (variant == 9) k_gru_fwd9<0><<<dim3(R, 2), HID>>>(fa); else if (variant == 10) k_gru_fwd9<0, 0><<<dim3(R, 2), HID>>>(fa); else if (variant == 11) k_gru_fwd11<0><<<dim3(R, 2), 256>>>(fa); else if (variant == 12) k_gru_fwd9<0><<<dim3(296, 1), HID>>>(fb); else k_gru_fwd8<0><<<dim3(R, 2), 128>>>(fa); };
    auto bwd = [&] { if (variant == 7) k_gru_bwd7<<<R, HID>>>(ba); else if (variant == 9 || variant == 10 || variant == 11 || variant == 12) k_gru_bwd9<<<R, HID>>>(ba); else k_gru_bwd8<<<R, 128>>>(ba); };
    const float fus = time_us(fwd);
    const float bus = time_us(bwd);
    printf("variant %d  fwd grid %dx2: %7.1f us  %5.0f cycles/step    bwd grid %d: %7.1f us  %5.0f cycles/step   (%s)\n", variant, R, fus,
           fus * 1965 / TT, R, bus, bus * 1965 / TT, cudaGetErrorString(cudaGetLastError()));

    // ---- fp64 reference on rows 0..NCHK-1 of the online net
    std::vector<float> hh(M * HID), gg(M * 256), dd(M * 256);
    cudaMemcpy(hh.data(), h[0], M * HID * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gg.data(), gates, M * 1024, cudaMemcpyDeviceToHost);
    cudaMemcpy(dd.data(), dg, M * 1024, cudaMemcpyDeviceToHost);
    const float *W = P0.data() + L.w_hh, *bh = P0.data() + L.b_hh;
    double eh = 0, ed = 0, mh = 0, md = 0;
    for (int row = 0; row < NCHK && row < R; ++row) {
        std::vector<double> hp(HID, 0.0);
        std::vector<std::vector<double>> Hs(TT, std::vector<double>(HID)), Rr(TT, Hs[0]), Zz(TT, Hs[0]), Nn(TT, Hs[0]), Gn(TT, Hs[0]);
        for (int t = 0; t < TT; ++t) {
            const float *g = gi0.data() + ((size_t)t * R + row) * G3;
            std::vector<double> hn(HID);
            for (int i = 0; i < HID; ++i) {
                double ar = bh[i], az = bh[HID + i], an = bh[2 * HID + i];
                for (int k = 0; k < HID; ++k) { ar += (double)W[i * HID + k] * hp[k]; az += (double)W[(HID + i) * HID + k] * hp[k]; an += (double)W[(2 * HID + i) * HID + k] * hp[k]; }
                const double r = 1 / (1 + exp(-(g[i] + ar))), z = 1 / (1 + exp(-(g[HID + i] + az))), n = tanh(g[2 * HID + i] + r * an);
                hn[i] = n + z * (hp[i] - n);
                Rr[t][i] = r; Zz[t][i] = z; Nn[t][i] = n; Gn[t][i] = an;
                const double got = hh[((size_t)t * R + row) * HID + i];
                eh = fmax(eh, fabs(got - hn[i])); mh = fmax(mh, fabs(hn[i]));
            }
            Hs[t] = hn; hp = hn;
        }
        std::vector<double> carry(HID, 0.0), dgh(G3, 0.0);
        for (int t = TT - 1; t >= 0; --t) {
            std::vector<double> ndgh(G3);
            for (int k = 0; k < HID; ++k) {
                double dh = carry[k] + (t < TT - 1 ? (double)dhh[((size_t)t * R + row) * HID + k] : 0.0);
                for (int j = 0; j < G3; ++j) dh += dgh[j] * (double)W[j * HID + k];
                const double r = Rr[t][k], z = Zz[t][k], n = Nn[t][k], gn = Gn[t][k], hprev = t > 0 ? Hs[t - 1][k] : 0.0;
                const double dn = dh * (1 - z), dz = dh * (hprev - n), dnp = dn * (1 - n * n), dzp = dz * z * (1 - z);
                const double drp = dnp * gn * r * (1 - r), dghn = dnp * r;
                carry[k] = dh * z;
                ndgh[k] = drp; ndgh[HID + k] = dzp; ndgh[2 * HID + k] = dghn;
                const float *got = dd.data() + ((size_t)t * R + row) * 256;
                const double ref[4] = {drp, dzp, dnp, dghn};
                for (int q = 0; q < 4; ++q) { ed = fmax(ed, fabs(got[q * HID + k] - ref[q])); md = fmax(md, fabs(ref[q])); }
            }
            dgh = ndgh;
        }
    }
    printf("vs fp64 (rows 0..%d): max |h err| %.2e (max |h| %.2e)   max |d_g err| %.2e (max |d_g| %.2e)\n", NCHK - 1, eh, mh, ed, md);
    return 0;
}
