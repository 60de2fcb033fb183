// `Homulator.run <configfile> <operationName> <maxExecutionLevel> <currentLevel> <alpha> [cluster] [--flags]`
// Drop-in for the reference's only executable (reference bench_test/bench_micro24.cpp:5-52): same
// positional arguments, same operation names, same config dump; but instead of building an instruction
// stream and clocking a pipeline model (reference src/Operation.cpp:1025-1112) it executes the operation on
// seeded synthetic data on the GPU and reports measured time, the reference-shaped instruction counts and
// the HBM roofline fraction.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "context.h"

using namespace hml;

__global__ void k_fill_uniform(u64 *out, size_t n_per_limb, int n_limbs_total, const u64 *limb_q, u64 seed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per_limb) return;
  const int limb = blockIdx.y;
  u64 z = seed + 0x9E3779B97F4A7C15ull * ((u64)limb * n_per_limb + i + 1);  // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  out[(size_t)limb * n_per_limb + i] = z % limb_q[limb];
  (void)n_limbs_total;
}

__global__ void k_flush(u64 *buf, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = i;
}

// Algorithmic traffic in limb-sized words W = 8N bytes (each stage reads each distinct input
// limb once and writes each output limb once, unfused).
static double ks_words(double L, double A) {
  const double E = L + A, beta = std::ceil(L / A);
  return 2 * L + 2 * L + beta * E + 2 * beta * E + (3 * beta + 2) * E + 4 * A + 4 * A + 2 * (A + L) + 4 * L + 6 * L;
}
static double op_words(const std::string &op, double L, double A) {
  if (op == "hmult") return ks_words(L, A) + 7 * L + 6 * L + 2 * (4 + 5 * (L - 1));
  if (op == "hrotate") return ks_words(L, A) + 4 * L + 3 * L;
  if (op == "pmult" || op == "padd") return 2 * 3 * L;  // plaintext limb is re-read for the second component
  return 2 * 3 * L;
}

// Measured HBM copy bandwidth of this pool's B200s: MEASURED_PEAKS.json (driver-written, at the repository root = one level
// above the directory this library lives in; HML_PEAKS_JSON overrides the path), else the profiling guide's fallback.
static double measured_hbm_gbs(std::string &source) {
  std::vector<std::string> cand;
  if (const char *e = getenv("HML_PEAKS_JSON")) cand.push_back(e);
  Dl_info info;
  if (dladdr((const void *)&measured_hbm_gbs, &info) && info.dli_fname) {
    std::string lib = info.dli_fname;
    const size_t sl = lib.rfind('/');
    if (sl != std::string::npos) cand.push_back(lib.substr(0, sl) + "/../MEASURED_PEAKS.json");
  }
  cand.push_back("MEASURED_PEAKS.json");
  for (const std::string &path : cand) {
    FILE *f = fopen(path.c_str(), "r");
    if (!f) continue;
    std::string text(1 << 14, '\0');
    text.resize(fread(&text[0], 1, text.size(), f));
    fclose(f);
    const size_t k = text.find("\"hbm_gbs\"");
    if (k == std::string::npos) continue;
    const size_t c = text.find(':', k);
    if (c == std::string::npos) continue;
    const double v = atof(text.c_str() + c + 1);
    if (v > 100.0) { source = "measured (MEASURED_PEAKS.json)"; return v; }
  }
  source = "fallback (B200_PROFILING.md)";
  return 6650.0;
}

static int fill(hml_ctx *ctx, u64 *dev, const std::vector<u64> &limb_mod, u64 seed) {
  u64 *dq = nullptr;
  if (!dev || cudaMalloc(&dq, limb_mod.size() * 8) != cudaSuccess) return 1;
  if (cudaMemcpy(dq, limb_mod.data(), limb_mod.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dq); return 1; }
  const size_t N = ctx->p.N;
  k_fill_uniform<<<dim3((unsigned)((N + 255) / 256), (unsigned)limb_mod.size()), 256>>>(dev, N, (int)limb_mod.size(), dq, seed);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(dq);
  return e == cudaSuccess ? 0 : 1;
}

extern "C" int hml_cli_main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "Usage: %s <configfile> <operationName> <maxExecutionLevel> <currentLevel> <alpha> [cluster]"
                    " [--iters K] [--warmup W] [--rot R] [--device D] [--no-flush] [--verify] [--packed-key]\n", argv[0]);
    return 1;
  }
  const std::string path = argv[1], op = argv[2];
  const uint32_t maxl = (uint32_t)std::atoi(argv[3]), L = (uint32_t)std::atoi(argv[4]), alpha = (uint32_t)std::atoi(argv[5]);
  int iters = 20, warmup = 3, rot = 1, device = 0, flush = 1, verify = 0, packed_key = 0, argi = 6;
  long cluster = -1;
  if (argi < argc && argv[argi][0] != '-') cluster = std::atol(argv[argi++]);  // accepted like the reference (bench_micro24.cpp:23-25)
  for (; argi < argc; ++argi) {
    const std::string f = argv[argi];
    auto val = [&](int &dst) { if (argi + 1 < argc) dst = std::atoi(argv[++argi]); };
    if (f == "--iters") val(iters);
    else if (f == "--warmup") val(warmup);
    else if (f == "--rot") val(rot);
    else if (f == "--device") val(device);
    else if (f == "--no-flush") flush = 0;
    else if (f == "--verify") verify = 1;
    else if (f == "--packed-key") packed_key = 1;  // the synthetic key goes through hml_key_pack (HML_KEY_PACKED)
  }
  CfgFile cfg;
  std::string err;
  if (!cfg.load(path, err)) {
    fprintf(stderr, "Error opening config file.\n%s\n", err.c_str());
    return 2;
  }
  fputs(cfg.dump().c_str(), stdout);  // dumped before the cluster override, like the reference
  if (!(op == "hmult" || op == "hrotate" || op == "hadd" || op == "pmult" || op == "padd")) {
    // reference bench_micro24.cpp:49-51: message on stdout, exit status 0
    printf("Error operation requirement, please double confirm!\n");
    return 0;
  }
  // the reference indexes its level tables without any check (and segfaults for hmult at L = 1); refuse instead
  if (maxl == 0 || alpha == 0 || L < (op == "hmult" ? 2u : 1u) || L > maxl) {
    fprintf(stderr, "homulator_b200: currentLevel %u out of range: %s needs %d <= currentLevel <= maxExecutionLevel = %u\n", L, op.c_str(),
            op == "hmult" ? 2 : 1, maxl);
    return 4;
  }
  hml_ctx *ctx = nullptr;
  int rc = hml_ctx_create(path.c_str(), maxl, alpha, device, &ctx);
  if (rc) {
    fprintf(stderr, "homulator_b200: cannot create context: %s\n", hml_last_create_error());
    return 3;
  }
  if (cluster >= 0) ctx->cfg.set("cluster", (uint32_t)cluster);
  hml_counts cnt;
  rc = hml_get_counts(ctx, op.c_str(), L, &cnt);
  if (rc) {
    fprintf(stderr, "homulator_b200: %s\n", hml_last_create_error());
    hml_ctx_destroy(ctx);
    return 4;
  }
  const Params &p = ctx->p;
  const size_t N = p.N;
  const uint32_t beta = p.beta(L), E = L + alpha;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  // ---- seeded synthetic operands: uniform residues per limb
  std::vector<u64> ct_mod, key_mod, pt_mod;
  for (int k = 0; k < 2; ++k) for (uint32_t i = 0; i < L; ++i) ct_mod.push_back(p.mod[i]);
  for (uint32_t i = 0; i < L; ++i) pt_mod.push_back(p.mod[i]);
  for (uint32_t j = 0; j < beta * 2; ++j) for (uint32_t e = 0; e < E; ++e) key_mod.push_back(p.mod[p.ext_mod(L, e)]);
  uint64_t *a = nullptr, *b = nullptr, *key = nullptr, *out = nullptr;
  u64 *flushbuf = nullptr;
  const u64 seed = 0x486F6D756C61746Full;
  int bad = cudaMalloc(&a, ct_mod.size() * N * 8) != cudaSuccess;
  bad |= cudaMalloc(&b, ct_mod.size() * N * 8) != cudaSuccess;
  bad |= cudaMalloc(&out, ct_mod.size() * N * 8) != cudaSuccess;
  bad |= fill(ctx, (u64 *)a, ct_mod, seed + 1);
  if (op == "pmult" || op == "padd") bad |= fill(ctx, (u64 *)b, pt_mod, seed + 2);
  else bad |= fill(ctx, (u64 *)b, ct_mod, seed + 2);
  if (op == "hmult" || op == "hrotate") {
    bad |= cudaMalloc(&key, key_mod.size() * N * 8) != cudaSuccess;
    bad |= fill(ctx, (u64 *)key, key_mod, seed + 3);
  }
  const size_t flush_words = (size_t)256 << 17;  // 256 MiB > L2
  if (flush) bad |= cudaMalloc(&flushbuf, flush_words * 8) != cudaSuccess;
  if (bad || cudaGetLastError() != cudaSuccess) {
    fprintf(stderr, "homulator_b200: device allocation / fill failed\n");
    return 5;
  }
  u64 g = 1;
  for (int r = 0; r < rot; ++r) g = (g * 5) % (2 * N);
  uint64_t *key_pk = nullptr;
  uint32_t key_limbs = L;
  if (packed_key && key && cluster < 2) {
    if (cudaMalloc(&key_pk, key_mod.size() * N * 8) != cudaSuccess || hml_key_pack(ctx, key, key_mod.size(), key_pk, nullptr) != HML_OK || hml_sync(ctx, nullptr) != HML_OK) {
      fprintf(stderr, "homulator_b200: key packing failed: %s\n", hml_last_error(ctx));
      return 5;
    }
    key_limbs |= HML_KEY_PACKED;
  }
  const uint64_t *key_used = key_pk ? key_pk : key;
  auto run = [&]() -> int {
    if (op == "hmult") return hml_hmult(ctx, L, a, b, key_used, key_limbs, out, nullptr);
    if (op == "hrotate") return hml_hrotate(ctx, L, a, key_used, key_limbs, g, out, nullptr);
    if (op == "hadd") return hml_hadd(ctx, L, a, b, out, nullptr);
    if (op == "pmult") return hml_pmult(ctx, L, a, b, out, nullptr);
    return hml_padd(ctx, L, a, b, out, nullptr);
  };
  {  // buffer plan, like the reference's `Malloc <name> from A to B` lines (reference include/Addr.h:46-47)
    std::vector<char> plan(8192);
    if (hml_buffer_plan(ctx, op.c_str(), L, plan.data(), plan.size()) == HML_OK) fputs(plan.data(), stdout);
  }
  std::string OP = op;
  std::transform(OP.begin(), OP.end(), OP.begin(), ::toupper);
  printf("\n\nWelcome! Start executing %s on %s (%d SMs)!\n\n", OP.c_str(), prop.name, prop.multiProcessorCount);
  time_t t0 = time(nullptr);
  printf("Start time: %s\n", ctime(&t0));
  for (int i = 0; i < warmup; ++i)
    if ((rc = run())) break;
  if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = HML_ERR_CUDA;
  if (rc) {
    fprintf(stderr, "homulator_b200: %s failed: %s\n", op.c_str(), hml_last_error(ctx));
    return 6;
  }
  hml_exec_counts_reset(ctx);
  run();
  hml_exec_counts ex;
  hml_exec_counts_get(ctx, &ex);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> us;
  for (int i = 0; i < iters; ++i) {
    if (flush) k_flush<<<1184, 256>>>(flushbuf, flush_words);
    cudaEventRecord(e0);
    run();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    us.push_back(ms * 1000.f);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    fprintf(stderr, "homulator_b200: CUDA failure during the timed runs\n");
    return 6;
  }
  std::sort(us.begin(), us.end());
  const double med = us.empty() ? 0 : us[us.size() / 2], mn = us.empty() ? 0 : us.front();
  const double words = op_words(op, L, alpha), bytes = words * 8.0 * N;
  const double gbs = med > 0 ? bytes / (med * 1e-6) / 1e9 : 0;
  std::string peak_src;
  const double peak = measured_hbm_gbs(peak_src);
  // one more run in measuring mode: device time per kernel class, the executed counterpart of the reference's per-unit
  // statistics dump (Statistic keys NTT_(c) / BCONV_(c) / EWE_(c) / AUTO_(c) / HBM_(c), reference include/Staistics.h:6-40)
  hml_profile prof{};
  if (hml_profile_begin(ctx, nullptr) == HML_OK) {
    run();
    if (hml_profile_end(ctx, &prof) != HML_OK) memset(&prof, 0, sizeof(prof));
  }
  time_t t1 = time(nullptr);
  printf("\n\nCompleted!\n");
  printf("%s executed\t%.2f us (median of %d, min %.2f us, L2 %s between runs)\n\n", OP.c_str(), med, iters, mn,
         flush ? "flushed" : "not flushed");
  printf("End time: %s\n", ctime(&t1));
  printf("Reference-trace instruction counts (one instruction = %u coefficients of one limb):\n", p.batch_size);
  printf("=====================================\n");
  printf("%-34s %-12s %10s %12s\n", "stage", "opcode", "limb-ops", "instructions");
  for (uint32_t i = 0; i < cnt.n_stages; ++i)
    printf("%-34s %-12s %10llu %12llu\n", cnt.stages[i].label, cnt.stages[i].opcode, (unsigned long long)cnt.stages[i].limb_ops,
           (unsigned long long)cnt.stages[i].instructions);
  printf("NTT :\t%llu\nINTT :\t%llu\nMULT :\t%llu\nBCONV_STEP2 :\t%llu\nAUTO :\t%llu\nTOTAL :\t%llu\ndriverTotal :\t%llu\n",
         (unsigned long long)cnt.ntt, (unsigned long long)cnt.intt, (unsigned long long)cnt.mult, (unsigned long long)cnt.bconv_step2,
         (unsigned long long)cnt.automorph, (unsigned long long)cnt.total, (unsigned long long)cnt.driver_total);
  printf("=====================================\n");
  printf("Executed on the GPU (limb-ops; deltas vs the trace are SURVEY.md 3.5 D1-D3):\n");
  printf("NTT_limbs :\t%llu\nINTT_limbs :\t%llu\nEWE_limbs :\t%llu\nBCONV_limb_MACs :\t%llu\nAUTO_limbs :\t%llu\nkernel_launches :\t%llu\n",
         (unsigned long long)ex.ntt_limbs, (unsigned long long)ex.intt_limbs, (unsigned long long)ex.ewe_limbs,
         (unsigned long long)ex.bconv_limb_macs, (unsigned long long)ex.automorph_limbs, (unsigned long long)ex.kernel_launches);
  printf("Algorithmic traffic :\t%.0f W = %.4e B -> %.1f GB/s = %.1f%% of the HBM peak (%.1f GB/s, %s)\n", words, bytes, gbs,
         100.0 * gbs / peak, peak, peak_src.c_str());
  {
    // per-class algorithmic bytes of the EXECUTED schedule: transforms and automorphisms 2 W per limb, conversions
    // (sources + targets) are folded into BCONV's limb-MAC count, element-wise passes ~3 W per output limb
    const double Wb = 8.0 * N;
    const double cls_bytes[HML_CLS_COUNT] = {2.0 * Wb * ex.ntt_limbs, 2.0 * Wb * ex.intt_limbs, 0.0, 3.0 * Wb * ex.ewe_limbs, 2.0 * Wb * ex.automorph_limbs};
    const char *names[HML_CLS_COUNT] = {"NTT", "INTT", "BCONV", "EWE", "AUTO"};
    printf("=====================================\n");
    printf("Per-unit statistics (cluster 0 = this GPU; measuring mode, kernels serialised):\n");
    for (int c = 0; c < HML_CLS_COUNT; ++c) {
      printf("%s_(0) :\tbusy %.2f us\t%llu launches", names[c], prof.us[c], (unsigned long long)prof.launches[c]);
      if (cls_bytes[c] > 0 && prof.us[c] > 0) printf("\t%.1f GB/s", cls_bytes[c] / (prof.us[c] * 1e-6) / 1e9);
      if (c == HML_CLS_BCONV && prof.us[c] > 0) printf("\t%.3e limb-MACs/s", (double)ex.bconv_limb_macs / (prof.us[c] * 1e-6));
      printf("\n");
    }
    printf("HBM_(0) :\t%.1f GB/s algorithmic over %.2f us serialised (%.1f%% of peak)\n", prof.total_us > 0 ? bytes / (prof.total_us * 1e-6) / 1e9 : 0.0,
           prof.total_us, prof.total_us > 0 ? 100.0 * bytes / (prof.total_us * 1e-6) / 1e9 / peak : 0.0);
  }
  // ---- [cluster] (reference bench_micro24.cpp:23-25 overrides the number of compute clusters, to which the reference maps
  // limb l as l % cluster, include/Driver.h:158,:178): here the cluster count is the limb-shard world.  With cluster >= 2 and
  // that many (or fewer: clipped) visible GPUs, hmult / hrotate run once more limb-sharded over the GPUs, one rank per device,
  // all driven from this process (hml_group_op), and the time is reported next to the one-GPU figure.
  int n_dev = 0;
  cudaGetDeviceCount(&n_dev);
  const int gpus = (cluster >= 2 && (op == "hmult" || op == "hrotate")) ? (int)std::min<long>(cluster, n_dev) : 1;
  double sharded_us = 0.0;
  int sharded_ok = -1;
  if (gpus >= 2) {
    const uint32_t W = (uint32_t)gpus, Lout = op == "hmult" ? L - 1 : L;
    std::vector<hml_ctx *> cx(W, nullptr);
    std::vector<hml_shard *> sh(W, nullptr);
    std::vector<uint64_t *> a_own(W, nullptr), b_own(W, nullptr), key_own(W, nullptr), out_own(W, nullptr);
    std::vector<cudaStream_t> st(W, nullptr);
    int src = 0;
    for (uint32_t r = 0; r < W && !src; ++r) {
      src = hml_ctx_create(path.c_str(), maxl, alpha, (device + (int)r) % n_dev, &cx[r]);
      if (!src) src = hml_shard_create(cx[r], L, r, W, &sh[r]);
    }
    if (!src) src = hml_shard_connect_local(sh.data(), W);
    for (uint32_t r = 0; r < W && !src; ++r) {
      src = hml_shard_prepare(sh[r], L);
      cudaSetDevice(cx[r]->device);
      cudaStreamCreateWithFlags(&st[r], cudaStreamNonBlocking);
      // this rank's limbs of the operands: ciphertexts [2][nq][N], key [beta][2][n_own_ext][N]
      std::vector<uint32_t> oq, oe;
      for (uint32_t i = r; i < L; i += W) oq.push_back(i);
      for (uint32_t e = r; e < E; e += W) oe.push_back(e);
      const size_t nq = oq.size(), ne = oe.size();
      cudaMalloc(&a_own[r], std::max<size_t>(1, 2 * nq) * N * 8);
      cudaMalloc(&b_own[r], std::max<size_t>(1, 2 * nq) * N * 8);
      cudaMalloc(&out_own[r], std::max<size_t>(1, 2 * nq) * N * 8);
      cudaMalloc(&key_own[r], std::max<size_t>(1, 2 * beta * ne) * N * 8);
      for (int c = 0; c < 2; ++c)
        for (size_t k = 0; k < nq; ++k) {
          cudaMemcpy(a_own[r] + (c * nq + k) * N, a + ((size_t)c * L + oq[k]) * N, N * 8, cudaMemcpyDefault);
          cudaMemcpy(b_own[r] + (c * nq + k) * N, b + ((size_t)c * L + oq[k]) * N, N * 8, cudaMemcpyDefault);
        }
      for (uint32_t jc = 0; jc < 2 * beta; ++jc)
        for (size_t k = 0; k < ne; ++k) cudaMemcpy(key_own[r] + (jc * ne + k) * N, key + ((size_t)jc * E + oe[k]) * N, N * 8, cudaMemcpyDefault);
      if (cudaDeviceSynchronize() != cudaSuccess) src = HML_ERR_CUDA;
    }
    const int kind = op == "hmult" ? 2 : 1;
    auto run_sharded = [&]() -> int {
      return hml_group_op(sh.data(), W, kind, L, (const uint64_t *const *)a_own.data(), (const uint64_t *const *)b_own.data(),
                          (const uint64_t *const *)key_own.data(), out_own.data(), nullptr, g, (void *const *)st.data());
    };
    auto sync_all = [&]() { for (uint32_t r = 0; r < W; ++r) { cudaSetDevice(cx[r]->device); cudaStreamSynchronize(st[r]); } };
    for (int i = 0; i < warmup + 1 && !src; ++i) src = run_sharded();
    sync_all();
    if (!src) {
      std::vector<double> ts;
      for (int i = 0; i < iters && !src; ++i) {
        const double t_0 = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }();
        src = run_sharded();
        sync_all();
        const double t_1 = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }();
        ts.push_back(t_1 - t_0);
      }
      std::sort(ts.begin(), ts.end());
      if (!ts.empty()) sharded_us = ts[ts.size() / 2];
    }
    for (uint32_t r = 0; r < W && !src; ++r) src = hml_shard_check(sh[r], st[r]);
    if (!src && verify) {  // the sharded result must equal the one-GPU result limb by limb
      cudaSetDevice(device);
      run();
      cudaDeviceSynchronize();
      std::vector<uint64_t> h1(N), h2(N);
      sharded_ok = 1;
      for (uint32_t r = 0; r < W && sharded_ok == 1; ++r) {
        std::vector<uint32_t> keep;
        for (uint32_t i = r; i < Lout; i += W) keep.push_back(i);
        for (int c = 0; c < 2 && sharded_ok == 1; ++c)
          for (size_t k = 0; k < keep.size(); ++k) {
            cudaMemcpy(h1.data(), out + ((size_t)c * Lout + keep[k]) * N, N * 8, cudaMemcpyDefault);
            cudaMemcpy(h2.data(), out_own[r] + (c * keep.size() + k) * N, N * 8, cudaMemcpyDefault);
            if (memcmp(h1.data(), h2.data(), N * 8) != 0) { sharded_ok = 0; break; }
          }
      }
    }
    if (src) fprintf(stderr, "homulator_b200: limb-sharded run over %d GPUs failed: %s\n", gpus, cx[0] ? hml_last_error(cx[0]) : hml_last_create_error());
    else printf("%s limb-sharded over %d GPUs (cluster = %ld)\t%.2f us (median of %d, host clock around all ranks)%s\n\n", OP.c_str(), gpus, cluster,
                sharded_us, iters, sharded_ok == 1 ? ", result identical to the one-GPU run" : sharded_ok == 0 ? ", RESULT DIFFERS from the one-GPU run" : "");
    for (uint32_t r = 0; r < W; ++r) {
      if (cx[r]) cudaSetDevice(cx[r]->device);
      cudaFree(a_own[r]); cudaFree(b_own[r]); cudaFree(key_own[r]); cudaFree(out_own[r]);
      if (st[r]) cudaStreamDestroy(st[r]);
      hml_shard_destroy(sh[r]);
      hml_ctx_destroy(cx[r]);
    }
    cudaSetDevice(device);
    if (src || sharded_ok == 0) { hml_ctx_destroy(ctx); return 7; }
  }
  printf("{\"op\": \"%s\", \"N\": %u, \"maxLevel\": %u, \"L\": %u, \"alpha\": %u, \"us_median\": %.3f, \"us_min\": %.3f, \"iters\": %d, "
         "\"l2_flushed\": %s, \"algorithmic_bytes\": %.0f, \"achieved_gbs\": %.2f, \"hbm_frac_of_measured\": %.4f, "
         "\"trace\": {\"NTT\": %llu, \"INTT\": %llu, \"MULT\": %llu, \"BCONV_STEP2\": %llu, \"AUTO\": %llu, \"total\": %llu, \"driverTotal\": %llu}, "
         "\"executed\": {\"ntt_limbs\": %llu, \"intt_limbs\": %llu, \"ewe_limbs\": %llu, \"bconv_limb_macs\": %llu, \"auto_limbs\": %llu, "
         "\"kernel_launches\": %llu}, \"class_us\": {\"NTT\": %.2f, \"INTT\": %.2f, \"BCONV\": %.2f, \"EWE\": %.2f, \"AUTO\": %.2f}, \"hbm_peak_gbs\": %.1f, "
         "\"hbm_peak_source\": \"%s\", \"cluster\": %ld, \"gpus_used\": %d, \"sharded_us_median\": %.3f, \"sharded_matches_one_gpu\": %s, \"gpu\": \"%s\"}\n",
         op.c_str(), p.N, maxl, L, alpha, med, mn, iters, flush ? "true" : "false", bytes, gbs, gbs / peak,
         (unsigned long long)cnt.ntt, (unsigned long long)cnt.intt, (unsigned long long)cnt.mult, (unsigned long long)cnt.bconv_step2,
         (unsigned long long)cnt.automorph, (unsigned long long)cnt.total, (unsigned long long)cnt.driver_total,
         (unsigned long long)ex.ntt_limbs, (unsigned long long)ex.intt_limbs, (unsigned long long)ex.ewe_limbs,
         (unsigned long long)ex.bconv_limb_macs, (unsigned long long)ex.automorph_limbs, (unsigned long long)ex.kernel_launches,
         prof.us[0], prof.us[1], prof.us[2], prof.us[3], prof.us[4], peak, peak_src.c_str(), cluster, gpus, sharded_us,
         sharded_ok == 1 ? "true" : sharded_ok == 0 ? "false" : "null", prop.name);
  cudaFree(a); cudaFree(b); cudaFree(out); cudaFree(key); cudaFree(key_pk); cudaFree(flushbuf);
  hml_ctx_destroy(ctx);
  return 0;
}
