// Stand-alone timing + cross-check harness for the GRU recurrence kernels of learner.cuh (not product code).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/gru_bench tools/gru_bench.cu
#include "../ma_league_b200/csrc/learner.cuh"
#include <vector>
#include <random>
#include <cstdlib>
void mal_set_error(const char *, ...) {}

static float *dev(const std::vector<float> &v) {
    float *p; cudaMalloc(&p, v.size() * 4); cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice); return p;
}
static double maxdiff(const float *a, const float *b, size_t n, double *ref_max) {
    std::vector<float> ha(n), hb(n);
    cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost);
    double m = 0, rm = 0;
    for (size_t i = 0; i < n; ++i) { m = fmax(m, fabs((double)ha[i] - hb[i])); rm = fmax(rm, fabs((double)hb[i])); }
    *ref_max = rm;
    return m;
}
static double maxdiff_gates(const float *v2, const float *v1, size_t M) {   // v2: [m][unit][4]   v1: [m][4][64]
    std::vector<float> a(M * 256), b(M * 256);
    cudaMemcpy(a.data(), v2, M * 1024, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), v1, M * 1024, cudaMemcpyDeviceToHost);
    double m = 0;
    for (size_t r = 0; r < M; ++r) for (int i = 0; i < 64; ++i) for (int g = 0; g < 4; ++g)
        m = fmax(m, fabs((double)a[r * 256 + i * 4 + g] - b[r * 256 + g * 64 + i]));
    return m;
}
template <typename F> static float time_us(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1000.f / reps;
}

int main(int argc, char **argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 160, TT = argc > 2 ? atoi(argv[2]) : 201;
    const int d_in = 64, A = 11;
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
    float *h[2][2], *gates[2], *dg[2];
    for (int v = 0; v < 2; ++v) {
        for (int n = 0; n < 2; ++n) { cudaMalloc(&h[v][n], M * HID * 4); cudaMemset(h[v][n], 0, M * HID * 4); }
        cudaMalloc(&gates[v], M * 256 * 4); cudaMemset(gates[v], 0, M * 256 * 4);
        cudaMalloc(&dg[v], M * 256 * 4); cudaMemset(dg[v], 0, M * 256 * 4);
    }
    auto fargs = [&](int v) {
        GruFwdArgs a; a.params[0] = dP0; a.params[1] = dP1; a.gi[0] = dgi0; a.gi[1] = dgi1; a.hout[0] = h[v][0]; a.hout[1] = h[v][1];
        a.gates = gates[v]; a.TT = TT; a.R = R; a.d_in = d_in; a.n_actions = A; return a;
    };
    auto bargs = [&](int v) {
        GruBwdArgs a; a.params = dP0; a.hout = h[0][0]; a.gates = gates[v]; a.dh_head = ddhh; a.d_g = dg[v]; a.TT = TT; a.R = R;
        a.d_in = d_in; a.n_actions = A; return a;
    };
    double rm;
#define FWD1(RT) { auto a = fargs(0); float us = time_us([&] { k_gru_fwd<RT><<<dim3((R + RT - 1) / RT, 2), 192>>>(a); }); \
    printf("fwd v1 RT=%d grid=%4d: %8.1f us  %6.0f cycles/step  (%s)\n", RT, 2 * ((R + RT - 1) / RT), us, us * 1965 / TT, cudaGetErrorString(cudaGetLastError())); }
#define FWD2(RT) { auto a = fargs(1); float us = time_us([&] { k_gru_fwd4<0><<<dim3(R, 2), 64>>>(a); }); \
    double d0 = maxdiff(h[1][0], h[0][0], M * HID, &rm), d1 = maxdiff(h[1][1], h[0][1], M * HID, &rm), d2 = maxdiff_gates(gates[1], gates[0], M); \
    printf("fwd v2 RT=%d grid=%4d: %8.1f us  %6.0f cycles/step  maxdiff h %.2e %.2e gates %.2e (%s)\n", RT, 2 * ((R + RT - 1) / RT), us, us * 1965 / TT, d0, d1, d2, cudaGetErrorString(cudaGetLastError())); }
#define FWD2D(RT, DBG) { auto a = fargs(1); float us = time_us([&] { k_gru_fwd4<DBG><<<dim3(R, 2), 64>>>(a); }); \
    double d0 = maxdiff(h[1][0], h[0][0], M * HID, &rm); printf("fwd v2 RT=%d DBG=%3d: %8.1f us  %6.0f cycles/step  maxdiff h %.2e (%s)\n", RT, DBG, us, us * 1965 / TT, d0, cudaGetErrorString(cudaGetLastError())); }
#define BWD1(RT) { auto a = bargs(0); float us = time_us([&] { k_gru_bwd<RT><<<(R + RT - 1) / RT, 192>>>(a); }); \
    printf("bwd v1 RT=%d grid=%4d: %8.1f us  %6.0f cycles/step  (%s)\n", RT, (R + RT - 1) / RT, us, us * 1965 / TT, cudaGetErrorString(cudaGetLastError())); }
#define BWD2(RT) { auto a = bargs(1); float us = time_us([&] { k_gru_bwd4<<<R, 64>>>(a); }); \
    double d0 = maxdiff(dg[1], dg[0], M * 256, &rm); \
    printf("bwd v2 RT=%d grid=%4d: %8.1f us  %6.0f cycles/step  maxdiff d_g %.2e (ref max %.2e) (%s)\n", RT, (R + RT - 1) / RT, us, us * 1965 / TT, d0, rm, cudaGetErrorString(cudaGetLastError())); }
    if (getenv("GRU_ONLY")) { FWD1(2); FWD2(1); BWD2(1); cudaDeviceSynchronize(); return 0; }
    FWD1(2); FWD1(4);
    FWD2(1);
    FWD2D(1, 1); FWD2D(1, 4); FWD2D(1, 8); FWD2D(1, 13); FWD2D(1, 128);
    FWD2(1);
    BWD1(1); BWD1(2);
    BWD2(1);
    cudaDeviceSynchronize();
    printf("final: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
