// test_f_update_mma2.cu -- standalone check + timing of csrc/f_update_mma2.cuh (pre-split fp16 factor + ldmatrix-fed mma.sync Gram
// kernel) against fp64 on the host and against the first-generation kernel of csrc/f_update_mma.cuh on the same inputs.
//
//   tools/test_f_update_mma2 [k] [small|c2|c5]
//
// small: ragged, badly scaled problem (empty series, tiny series, columns of very different magnitude), every series checked.
// c2 / c5: BASELINE-shaped sizes (T = n = 10 000, p = 0.9, k = 40 / T = 100 000, n = 40 000, p = 0.02, k = 64), sampled check + timing.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DTRMF_F32=1 -DValueType=float -o tools/test_f_update_mma2 tools/test_f_update_mma2.cu
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <time.h>
#include <unistd.h>

#include "../exp-trmf-nips16_b200/csrc/f_update_mma.cuh"
#include "../exp-trmf-nips16_b200/csrc/f_update_mma2.cuh"

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

    float *dXh;   // the pre-split copy the new kernel gathers from
    uint32_t *dyh;
    CHECK(cudaMalloc(&dyh, std::max<size_t>(nnz, 1) * 4));
    CHECK(cudaMalloc(&dXh, T * fm::Cfg2<K>::xh_floats_per_row * 4));
    {
        cudaFuncAttributes fa;
        const bool wide = n < (uint32_t)(24 * sms);
        printf("  (%s CTAs)\n", wide ? "wide" : "narrow");
        (void)fa;
    }
    float ms_d = 0, ms_g = 0, ms_md = 0, ms_mg = 0;
    auto watchdog = [&](cudaEvent_t ev, const char *what) {
        for (int i = 0; i < 1500; ++i) {
            if (cudaEventQuery(ev) == cudaSuccess) return;
            struct timespec ts = {0, 10 * 1000 * 1000};
            nanosleep(&ts, nullptr);
        }
        printf("HANG in %s\n", what);
        fflush(stdout);
        _exit(3);
    };
    unsigned long long launches = 0;
    // ---------------- MODE_DEFER ----------------
    for (int rep = 0; rep < 3; ++rep) {
        CHECK(cudaMemset(dsys, 0, (size_t)n * sysd * 8));
        CHECK(cudaEventRecord(e0));
        if (rep == 0 || getenv("FM2_SPLIT_EVERY_TIME")) { if (f_update_mma2_split_y(nullptr, sms, dval, dptr, n, dyh, dysc, dq + 300, &launches)) { printf("split_y failed\n"); return 1; } }
        if (f_update_mma2_launch<fm::MODE_DEFER>(nullptr, sms, dptr, didx, reinterpret_cast<const float *>(dyh), dX, T, dXh, dinvs, dF, (float *)nullptr, K, 0.5, n, dq, &launches, nullptr, 0, dsys, true, dysc)) { printf("mma2 launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        CHECK(cudaEventRecord(e1));
        watchdog(e1, "MODE_DEFER");
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        CHECK(cudaEventElapsedTime(&ms_d, e0, e1));
    }
    std::vector<double> sys_tc((size_t)n * sysd);
    CHECK(cudaMemcpy(sys_tc.data(), dsys, sys_tc.size() * 8, cudaMemcpyDeviceToHost));
    // the mma.sync kernel on the same inputs
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
    printf("  MODE_DEFER    mma2: %8.3f ms = %6.2f G entries/s | worst Gram rel. error %.2e, rhs %.2e  (%zu series checked)\n", ms_d, nnz / (ms_d * 1e-3) / 1e9, wG, wR, sample.size());
    printf("              mma.sync: %7.3f ms = %6.2f G entries/s | worst Gram rel. error %.2e, rhs %.2e\n", ms_md, nnz / (ms_md * 1e-3) / 1e9, mG, mR);

    // ---------------- MODE_GRAD ----------------
    std::vector<float> F0((size_t)n * K);
    for (auto &v : F0) v = (float)(rnd() - 0.5);
    std::vector<float> Gtc((size_t)n * K * K), Ftc((size_t)n * K), Gmm((size_t)n * K * K), Fmm((size_t)n * K);
    std::vector<double> frtc(n), frmm(n);
    for (int rep = 0; rep < 2; ++rep) {
        CHECK(cudaMemcpy(dF, F0.data(), F0.size() * 4, cudaMemcpyHostToDevice));
        CHECK(cudaEventRecord(e0));
        if (f_update_mma2_launch<fm::MODE_GRAD>(nullptr, sms, dptr, didx, dval, dX, T, dXh, dinvs, dF, dG, K, 0.0, n, dq, &launches, dW, 1, dfrow)) { printf("mma2 launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
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
    printf("  MODE_GRAD     mma2: %8.3f ms = %6.2f G entries/s | worst Gram %.2e, gradient row %.2e, loss %.2e\n", ms_g, nnz / (ms_g * 1e-3) / 1e9, eG[0], eF[0], eL[0]);
    printf("              mma.sync: %7.3f ms = %6.2f G entries/s | worst Gram %.2e, gradient row %.2e, loss %.2e\n", ms_mg, nnz / (ms_mg * 1e-3) / 1e9, eG[1], eF[1], eL[1]);
    cudaFree(dptr); cudaFree(didx); cudaFree(dval); cudaFree(dX); cudaFree(dXs); cudaFree(dinvs); cudaFree(dF); cudaFree(dG); cudaFree(dW);
    cudaFree(dq); cudaFree(dsys); cudaFree(dfrow); cudaFree(dysc);
    return 0;
}

int main(int argc, char **argv) {
    const int k = argc > 1 ? atoi(argv[1]) : 40;
    const char *size = argc > 2 ? argv[2] : "small";
    switch (k) {
        case 8: return run<8>(size);
        case 20: return run<20>(size);
        case 24: return run<24>(size);
        case 40: return run<40>(size);
        case 48: return run<48>(size);
        case 56: return run<56>(size);
        case 60: return run<60>(size);
        case 64: return run<64>(size);
    }
    printf("k must be one of 8 20 24 40 48 56 60 64\n");
    return 1;
}
