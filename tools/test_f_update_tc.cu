// test_f_update_tc.cu -- standalone check + timing of csrc/f_update_tc.cuh (the tcgen05 Gram kernel) against fp64 on the host
// and against the mma.sync kernel of csrc/f_update_mma.cuh on the same inputs.
//
//   tools/test_f_update_tc [k=40|64] [small|c2|c5]
//
// small: ragged, badly scaled problem (empty series, tiny series, columns of very different magnitude), every series checked.
// c2 / c5: BASELINE-shaped sizes (T = n = 10 000, p = 0.9, k = 40 / T = 100 000, n = 40 000, p = 0.02, k = 64), sampled check + timing.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTRMF_F32=1 -DValueType=float -o tools/test_f_update_tc tools/test_f_update_tc.cu
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <time.h>
#include <unistd.h>

#define TC_DEBUG 1
#include "../exp-trmf-nips16_b200/csrc/f_update_mma.cuh"
#include "../exp-trmf-nips16_b200/csrc/f_update_tc.cuh"

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static double relerr(const double *a, const double *b, size_t n) {
    double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) { const double d = a[i] - b[i]; num += d * d; den += b[i] * b[i]; }
    return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

template <int K>
static int run(const char *size) {
    const bool small = !strcmp(size, "small"), c5 = !strcmp(size, "c5");
    const size_t T = small ? 3000 : (c5 ? 100000 : 10000);
    const uint32_t n = small ? 700 : (c5 ? 40000 : 10000);
    const double p = small ? 0.6 : (c5 ? 0.02 : 0.9);
    srand(11);
    auto rnd = []() { return (double)rand() / RAND_MAX; };
    std::vector<float> X(T * K), Wv((size_t)n * K);
    for (size_t i = 0; i < T; ++i)
        for (int c = 0; c < K; ++c)
            X[i * K + c] = small ? (float)((rnd() - 0.3) * (c % 7 == 0 ? 40.0 : 1.0) * (c == 3 ? 1e-3 : 1.0)) : (float)rnd();
    for (auto &v : Wv) v = (float)(rnd() * 0.5);
    std::vector<uint64_t> ptr(n + 1, 0);
    std::vector<uint32_t> idx;
    std::vector<float> val;
    idx.reserve((size_t)(T * n * p * 1.05));
    val.reserve(idx.capacity());
    for (uint32_t j = 0; j < n; ++j) {
        double dens = p;
        if (small) dens = j == 5 ? 0.0 : (j % 97 == 0 ? 0.002 : (j % 13 == 0 ? 0.05 : 0.4 + 0.5 * rnd()));
        // geometric skipping keeps the generator O(nnz)
        if (dens > 0) {
            const double lq = std::log(1.0 - dens);
            for (double i = std::floor(std::log(1.0 - rnd() * 0.999999) / lq); i < (double)T; i += 1.0 + std::floor(std::log(1.0 - rnd() * 0.999999) / lq)) {
                idx.push_back((uint32_t)i);
                val.push_back((float)(rnd() * 6 - 2));
            }
        }
        ptr[j + 1] = idx.size();
    }
    const size_t nnz = idx.size();
    printf("k = %d, %s: T = %zu, n = %u, nnz = %zu\n", K, size, T, n, nnz);
    fflush(stdout);

    uint64_t *dptr; uint32_t *didx; float *dval, *dX, *dXs, *dinvs, *dF, *dG, *dW, *dysc; unsigned *dq; double *dsys, *dfrow;
    const size_t ld = K + 1, sysd = (K + 1) * ld;
    CHECK(cudaMalloc(&dptr, (n + 1) * 8)); CHECK(cudaMalloc(&didx, std::max<size_t>(nnz, 1) * 4)); CHECK(cudaMalloc(&dval, std::max<size_t>(nnz, 1) * 4));
    CHECK(cudaMalloc(&dX, X.size() * 4)); CHECK(cudaMalloc(&dXs, X.size() * 4)); CHECK(cudaMalloc(&dinvs, 128 * 4));
    CHECK(cudaMalloc(&dF, (size_t)n * K * 4)); CHECK(cudaMalloc(&dG, (size_t)n * K * K * 4)); CHECK(cudaMalloc(&dW, Wv.size() * 4));
    CHECK(cudaMalloc(&dq, 1024 * 4)); CHECK(cudaMalloc(&dsys, (size_t)n * sysd * 8)); CHECK(cudaMalloc(&dfrow, (size_t)n * 8));
    CHECK(cudaMalloc(&dysc, 2 * 4));
    CHECK(cudaMemcpy(dptr, ptr.data(), (n + 1) * 8, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(didx, idx.data(), nnz * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dval, val.data(), nnz * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dW, Wv.data(), Wv.size() * 4, cudaMemcpyHostToDevice));
    int dev = 0, sms = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));

    // column scaling (as f_update_mma_launch does), weight scale
    CHECK(cudaMemset(dq, 0, 1024 * 4));
    fm::colscale_max_kernel<<<4 * sms, 256>>>(dX, T, K, dq + 8);
    fm::colscale_apply_kernel<<<4 * sms, 256, sizeof(float) * K>>>(dX, T, K, dq + 8, dXs, dinvs);
    fm::colscale_max_kernel<<<4 * sms, 256>>>(dW, n, K, dq + 256);
    unsigned char *dXh;
    CHECK(cudaMalloc(&dXh, T * (size_t)tc::Cfg<K>::ROWB));
    tc::presplit_kernel<<<4 * sms, 256, sizeof(float) * K>>>(dX, T, K, tc::Cfg<K>::NG, dq + 8, reinterpret_cast<__half *>(dXh), dinvs);
    CHECK(cudaGetLastError());

    auto ktc_d = tc::f_update_tc_kernel<K, tc::MODE_DEFER>;
    auto ktc_g = tc::f_update_tc_kernel<K, tc::MODE_GRAD>;
    const size_t smem = tc::Cfg<K>::smem;
    CHECK(cudaFuncSetAttribute(ktc_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CHECK(cudaFuncSetAttribute(ktc_g, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = std::min<unsigned>(sms, n);
    {   // setmaxnreg's arithmetic in MODE_GRAD assumes ptxas' launch allocation: verify before launching
        cudaFuncAttributes fa;
        CHECK(cudaFuncGetAttributes(&fa, ktc_g));
        printf("  MODE_GRAD kernel: %d registers at launch (expected %d), drain warps raise to %d\n", fa.numRegs, tc::Cfg<K>::launch_regs(true), tc::Cfg<K>::drain_regs(true));
        if (fa.numRegs != tc::Cfg<K>::launch_regs(true)) { printf("register count mismatch: not launching\n"); return 1; }
        CHECK(cudaFuncGetAttributes(&fa, ktc_d));
        printf("  MODE_DEFER kernel: %d registers at launch (expected %d), drain warps raise to %d\n", fa.numRegs, tc::Cfg<K>::launch_regs(false), tc::Cfg<K>::drain_regs(false));
        if (fa.numRegs != tc::Cfg<K>::launch_regs(false)) { printf("register count mismatch: not launching\n"); return 1; }
    }
    float ms_d = 0, ms_g = 0, ms_md = 0, ms_mg = 0;

    // progress words in host-mapped memory + a watchdog: a hung kernel is reported (per-warp progress of CTA 0), not waited for
    unsigned *hdbg = nullptr, *ddbg = nullptr;
    CHECK(cudaHostAlloc(&hdbg, 32 * 4 * sizeof(unsigned), cudaHostAllocMapped));
    memset(hdbg, 0, 32 * 4 * sizeof(unsigned));
    CHECK(cudaHostGetDevicePointer(&ddbg, hdbg, 0));
    if (getenv("TC_DBG")) CHECK(cudaMemcpyToSymbol(tc::g_tc_dbg, &ddbg, sizeof ddbg));   // (slows CTA 0 down: diagnosis only)
    auto watchdog = [&](cudaEvent_t ev, const char *what) {
        for (int i = 0; i < 500; ++i) {
            if (cudaEventQuery(ev) == cudaSuccess) return;
            struct timespec ts = {0, 10 * 1000 * 1000};
            nanosleep(&ts, nullptr);
        }
        printf("HANG in %s; progress of CTA 0 (warp: setmaxnreg stage, waiting-for chunk+1, passed, extra):\n", what);
        for (int w = 0; w < 26; ++w) printf("  warp %2d: %u %u %u %u\n", w, hdbg[w * 4], hdbg[w * 4 + 1], hdbg[w * 4 + 2], hdbg[w * 4 + 3]);
        fflush(stdout);
        _exit(3);
    };
    // ---------------- MODE_DEFER ----------------
    for (int rep = 0; rep < 3; ++rep) {
        CHECK(cudaMemset(dsys, 0, (size_t)n * sysd * 8));
        CHECK(cudaEventRecord(e0));
        CHECK(cudaMemsetAsync(dq + 128, 0, 4));
        tc::absmax_range_kernel<<<2 * sms, 256>>>(dval, dptr, n, dq + 128);
        tc::weight_scale_kernel<<<1, 1>>>(dq + 128, nullptr, nullptr, K, dysc);
        ktc_d<<<grid, 32 * tc::Cfg<K>::nwarps(false), smem>>>(dptr, didx, dval, dXh, dinvs, dF, nullptr, n, nullptr, 0, dsys, dysc);
        CHECK(cudaEventRecord(e1));
        watchdog(e1, "MODE_DEFER");
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        CHECK(cudaEventElapsedTime(&ms_d, e0, e1));
    }
    std::vector<double> sys_tc((size_t)n * sysd);
    CHECK(cudaMemcpy(sys_tc.data(), dsys, sys_tc.size() * 8, cudaMemcpyDeviceToHost));
    // the mma.sync kernel on the same inputs
    unsigned long long launches = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CHECK(cudaMemset(dsys, 0, (size_t)n * sysd * 8));
        CHECK(cudaEventRecord(e0));
        if (f_update_mma_launch<fm::MODE_DEFER>(nullptr, sms, dptr, didx, dval, dX, T, dXs, dinvs, dF, (float *)nullptr, K, 0.5, n, dq, &launches, nullptr, 0, dsys)) { printf("mma launch failed\n"); return 1; }
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        CHECK(cudaEventElapsedTime(&ms_md, e0, e1));
    }
    std::vector<double> sys_mma((size_t)n * sysd);
    CHECK(cudaMemcpy(sys_mma.data(), dsys, sys_mma.size() * 8, cudaMemcpyDeviceToHost));

    // fp64 on the host for a sample of series
    std::vector<uint32_t> sample;
    if (small) for (uint32_t j = 0; j < n; ++j) sample.push_back(j);
    else for (uint32_t j : {0u, 1u, 2u, 77u, 1000u, n / 2, n - 2, n - 1}) sample.push_back(j);
    double wG = 0, wR = 0, mG = 0, mR = 0;
    std::vector<double> G((size_t)K * K), R(K), g2((size_t)K * K), r2(K);
    for (uint32_t j : sample) {
        std::fill(G.begin(), G.end(), 0.0); std::fill(R.begin(), R.end(), 0.0);
        for (uint64_t e = ptr[j]; e < ptr[j + 1]; ++e) {
            const float *x = &X[(size_t)idx[e] * K];
            for (int a = 0; a < K; ++a) {
                R[a] += (double)val[e] * (double)x[a];
                for (int b = 0; b <= a; ++b) G[(size_t)a * K + b] += (double)x[a] * (double)x[b];
            }
        }
        if (ptr[j + 1] == ptr[j]) continue;
        for (int which = 0; which < 2; ++which) {
            const double *s = (which ? sys_mma.data() : sys_tc.data()) + (size_t)j * sysd;
            for (int a = 0; a < K; ++a) { r2[a] = s[K * ld + a]; for (int b = 0; b < K; ++b) g2[(size_t)a * K + b] = b <= a ? s[a * ld + b] : 0.0; }
            const double eg = relerr(g2.data(), G.data(), (size_t)K * K), er = relerr(r2.data(), R.data(), K);
            if (which) { mG = std::max(mG, eg); mR = std::max(mR, er); } else { wG = std::max(wG, eg); wR = std::max(wR, er); }
            if (!which && (eg > 2e-6 || er > 2e-6 || eg != eg)) printf("  series %u (%llu entries): Gram %.2e rhs %.2e  <-- BAD\n", j, (unsigned long long)(ptr[j + 1] - ptr[j]), eg, er);
        }
    }
    auto dump_clk = [&](const char *what, int nw) {
        unsigned long long h[32 * 8];
        CHECK(cudaMemcpyFromSymbol(h, tc::g_tc_clk, sizeof h));
        printf("  %s: CTA 0 cycles per warp [wait A, wait/section B, section C, .., whole role]\n", what);
        for (int w = 0; w < nw; ++w) printf("    warp %2d: %10llu %10llu %10llu | %10llu\n", w, h[w * 8], h[w * 8 + 1], h[w * 8 + 2], h[w * 8 + 7]);
    };
    if (getenv("TC_CLK")) dump_clk("MODE_DEFER (drain: wait afull, epilogue | producer: wait empty, wait_group, issue | mma: wait full, wait aempty, issue)", tc::Cfg<K>::nwarps(false));
    printf("  MODE_DEFER  tcgen05: %8.3f ms = %6.2f G entries/s | worst Gram rel. error %.2e, rhs %.2e  (%zu series checked)\n", ms_d, nnz / (ms_d * 1e-3) / 1e9, wG, wR, sample.size());
    printf("              mma.sync: %7.3f ms = %6.2f G entries/s | worst Gram rel. error %.2e, rhs %.2e\n", ms_md, nnz / (ms_md * 1e-3) / 1e9, mG, mR);

    // ---------------- MODE_GRAD ----------------
    std::vector<float> F0((size_t)n * K);
    for (auto &v : F0) v = (float)(rnd() - 0.5);
    std::vector<float> Gtc((size_t)n * K * K), Ftc((size_t)n * K), Gmm((size_t)n * K * K), Fmm((size_t)n * K);
    std::vector<double> frtc(n), frmm(n);
    for (int rep = 0; rep < 2; ++rep) {
        CHECK(cudaMemcpy(dF, F0.data(), F0.size() * 4, cudaMemcpyHostToDevice));
        CHECK(cudaEventRecord(e0));
        CHECK(cudaMemsetAsync(dq + 128, 0, 4));
        tc::absmax_range_kernel<<<2 * sms, 256>>>(dval, dptr, n, dq + 128);
        tc::weight_scale_kernel<<<1, 1>>>(dq + 128, dq + 8, dq + 256, K, dysc);
        ktc_g<<<grid, 32 * tc::Cfg<K>::nwarps(true), smem>>>(dptr, didx, dval, dXh, dinvs, dF, dG, n, dW, 1, dfrow, dysc);
        CHECK(cudaEventRecord(e1));
        watchdog(e1, "MODE_GRAD");
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        CHECK(cudaEventElapsedTime(&ms_g, e0, e1));
    }
    CHECK(cudaMemcpy(Gtc.data(), dG, Gtc.size() * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(Ftc.data(), dF, Ftc.size() * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(frtc.data(), dfrow, n * 8, cudaMemcpyDeviceToHost));
    for (int rep = 0; rep < 2; ++rep) {
        CHECK(cudaMemcpy(dF, F0.data(), F0.size() * 4, cudaMemcpyHostToDevice));
        CHECK(cudaEventRecord(e0));
        if (f_update_mma_launch<fm::MODE_GRAD>(nullptr, sms, dptr, didx, dval, dX, T, dXs, dinvs, dF, dG, K, 0.0, n, dq, &launches, dW, 1, dfrow)) { printf("mma launch failed\n"); return 1; }
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        CHECK(cudaEventElapsedTime(&ms_mg, e0, e1));
    }
    CHECK(cudaMemcpy(Gmm.data(), dG, Gmm.size() * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(Fmm.data(), dF, Fmm.size() * 4, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(frmm.data(), dfrow, n * 8, cudaMemcpyDeviceToHost));
    double eG[2] = {0, 0}, eF[2] = {0, 0}, eL[2] = {0, 0};
    std::vector<double> Gf((size_t)K * K), grad(K), gg((size_t)K * K), gr(K);
    for (uint32_t j : sample) {
        std::fill(Gf.begin(), Gf.end(), 0.0);
        double loss = 0;
        for (int a = 0; a < K; ++a) grad[a] = (double)F0[(size_t)j * K + a];
        for (uint64_t e = ptr[j]; e < ptr[j + 1]; ++e) {
            const float *x = &X[(size_t)idx[e] * K];
            double z = -(double)val[e];
            for (int a = 0; a < K; ++a) z += (double)Wv[(size_t)j * K + a] * (double)x[a];
            loss += z * z;
            for (int a = 0; a < K; ++a) {
                grad[a] += z * (double)x[a];
                for (int b = 0; b < K; ++b) Gf[(size_t)a * K + b] += (double)x[a] * (double)x[b];
            }
        }
        for (int which = 0; which < 2; ++which) {
            const float *gs = (which ? Gmm.data() : Gtc.data()) + (size_t)j * K * K, *fs = (which ? Fmm.data() : Ftc.data()) + (size_t)j * K;
            const double fl = which ? frmm[j] : frtc[j];
            for (int q = 0; q < K * K; ++q) gg[q] = gs[q];
            for (int q = 0; q < K; ++q) gr[q] = fs[q];
            const double a = relerr(gg.data(), Gf.data(), (size_t)K * K), b = relerr(gr.data(), grad.data(), K);
            const double c = loss > 0 ? std::fabs(fl - loss) / loss : std::fabs(fl);
            eG[which] = std::max(eG[which], a); eF[which] = std::max(eF[which], b); eL[which] = std::max(eL[which], c);
            if (!which && (a > 2e-6 || b > 1e-4 || c > 1e-5 || a != a || b != b)) printf("  series %u (%llu entries): Gram %.2e grad %.2e loss %.2e  <-- BAD\n", j, (unsigned long long)(ptr[j + 1] - ptr[j]), a, b, c);
        }
    }
    if (getenv("TC_CLK")) dump_clk("MODE_GRAD (+ residual: wait full, compute)", tc::Cfg<K>::nwarps(true));
    printf("  MODE_GRAD   tcgen05: %8.3f ms = %6.2f G entries/s | worst Gram %.2e, gradient row %.2e, loss %.2e\n", ms_g, nnz / (ms_g * 1e-3) / 1e9, eG[0], eF[0], eL[0]);
    printf("              mma.sync: %7.3f ms = %6.2f G entries/s | worst Gram %.2e, gradient row %.2e, loss %.2e\n", ms_mg, nnz / (ms_mg * 1e-3) / 1e9, eG[1], eF[1], eL[1]);
    cudaFree(dptr); cudaFree(didx); cudaFree(dval); cudaFree(dX); cudaFree(dXs); cudaFree(dinvs); cudaFree(dF); cudaFree(dG); cudaFree(dW);
    cudaFree(dq); cudaFree(dsys); cudaFree(dfrow); cudaFree(dysc);
    return 0;
}

int main(int argc, char **argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 40;
    const char *size = argc > 2 ? argv[2] : "small";
    if (k == 40) return run<40>(size);
    if (k == 64) return run<64>(size);
    printf("k must be 40 or 64\n");
    return 1;
}
