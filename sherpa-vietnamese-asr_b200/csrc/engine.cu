// Engine: weight container loading, workspace management, the fbank -> encoder -> search pipeline over a
// ragged batch, and the C-ABI declared in include/b200asr.h.
// The pipeline stands where the reference runs compute_fbank_ort + `_ort_beam_search` per chunk
// (/root/reference core/asr_engine.py:1209-1253, chunk loop :2326-2397); the encoder schedule follows
// SURVEY.md Appendix B (icefall Zipformer2 inference graph).
#include <cuda.h>
#include <math.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200asr.h"
#include "common.cuh"
#include "context_graph.h"

namespace b200asr {

void launch_gemm_tc(const GemmArgs &g, cudaStream_t st);    // gemm_tc.cu: tcgen05, TF32 operands
void launch_gemm_tc3(const GemmArgs &g, cudaStream_t st);   // gemm_tc.cu: fp32-grade on tcgen05 (fp16 operand split, or 3xTF32)
void launch_gemm_bf16(const GemmArgs &g, cudaStream_t st);  // gemm_tc.cu: BF16 operands, FP32 accumulate
bool gemm_tc_available();

namespace {

thread_local std::string g_last_error;

// ------------------------------------------------------------------ growable device buffer
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  template <typename T>
  T *get(size_t n) {
    const size_t need = n * sizeof(T);
    if (need > cap) {
      if (p) CUDA_CHECK(cudaFree(p));
      p = nullptr;
      cap = need + need / 8 + 256;
      CUDA_CHECK(cudaMalloc(&p, cap));
    }
    return reinterpret_cast<T *>(p);
  }
  template <typename T>
  T *ptr() const { return reinterpret_cast<T *>(p); }
};

struct Tensor {
  std::vector<int> shape;
  std::vector<float> host;
  float *dev = nullptr;
  size_t numel() const { size_t n = 1; for (int d : shape) n *= (size_t)d; return n; }
};

std::vector<int> parse_int_list(const std::string &s) {
  std::vector<int> v;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ',')) if (!tok.empty()) v.push_back(atoi(tok.c_str()));
  return v;
}

struct LayerW {
  const float *attn_in_w, *attn_in_b, *pos_w;
  const float *ff_in_w[3], *ff_in_b[3], *ff_out_w[3], *ff_out_b[3];
  int ff_dim[3];
  const float *nl_in_w, *nl_in_b, *nl_out_w, *nl_out_b;
  const float *sa_in_w[2], *sa_in_b[2], *sa_out_w[2], *sa_out_b[2];
  const float *cv_in_w[2], *cv_in_b[2], *cv_dw_w[2], *cv_dw_b[2], *cv_out_w[2], *cv_out_b[2];
  const float *norm_bias, *norm_log_scale, *bypass, *bypass_mid;
};

struct StackW {
  int L, ds, D, F, H, k;
  const float *ds_bias = nullptr, *combiner = nullptr;
  std::vector<LayerW> layers;
};

struct Timings { float fbank = 0, encoder = 0, search = 0, total = 0, h2d = 0, d2h = 0; };

// One search in flight: its own stream (high priority), search state and encoder_out buffer. A decode pass splits its batch
// into length-sorted groups; group g's search runs on lane g while the main stream already encodes group g + 1.
constexpr int kMaxLanes = 8;
struct Lane {
  cudaStream_t st = nullptr;
  SearchState *search = nullptr;
  DevBuf enc, soff;
  cudaEvent_t f0 = nullptr, f1 = nullptr, e1 = nullptr, s0 = nullptr, s1 = nullptr;
  std::vector<int> members;   // original utterance indices of the group this lane holds
  std::vector<int> Tp;
  int steps = 0;
};

}  // namespace

struct Stream;

// Driver entry points of the green-context API, resolved at run time (the library must load on machines without libcuda,
// e.g. the CPU-only build check).
namespace {
struct GreenApi {
  CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
  CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
  CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
  CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
  bool ok = false;
  GreenApi() {
    auto get = [](const char *name, void **fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && *fn && q == cudaDriverEntryPointSuccess;
    };
    ok = get("cuDeviceGet", (void **)&DeviceGet) && get("cuDeviceGetDevResource", (void **)&DeviceGetDevResource) &&
         get("cuDevSmResourceSplitByCount", (void **)&DevSmResourceSplitByCount) &&
         get("cuDevResourceGenerateDesc", (void **)&DevResourceGenerateDesc) && get("cuGreenCtxCreate", (void **)&GreenCtxCreate) &&
         get("cuGreenCtxDestroy", (void **)&GreenCtxDestroy) && get("cuGreenCtxStreamCreate", (void **)&GreenCtxStreamCreate);
    if (!ok) cudaGetLastError();
  }
};
const GreenApi &green_api() { static GreenApi a; return a; }
}  // namespace


// Page-locked host memory for stream PCM: accept_waveform copies the caller's samples straight into pinned
// memory, so decode_streams issues true asynchronous H2D copies at PCIe speed. cudaHostAlloc costs ~2 ms per call,
// far too slow per stream, so blocks (power-of-two size classes) are carved from 64 MB pinned slabs and recycled
// through per-class free lists; slabs live for the life of the process. Past kMaxPinned the pool hands out
// pageable blocks (cudaMemcpyAsync stages those itself) instead of pinning unbounded host memory.
class PinnedPool {
 public:
  static PinnedPool &get() { static PinnedPool p; return p; }
  void *alloc(size_t bytes, size_t *cap) {
    size_t c = 64 * 1024;
    while (c < bytes) c <<= 1;
    *cap = c;
    std::lock_guard<std::mutex> lk(mu_);
    auto &fl = free_[c];
    if (!fl.empty()) { void *p = fl.back(); fl.pop_back(); return p; }
    if (c > slab_left_) {
      const size_t slab = std::max(c, kSlab);
      void *p = nullptr;
      if (pinned_total_ + slab > kMaxPinned || cudaHostAlloc(&p, slab, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        p = malloc(c);   // pageable block; still recycled through the free list
        if (!p) throw std::runtime_error("out of host memory for stream audio");
        return p;
      }
      pinned_total_ += slab;
      slab_cur_ = reinterpret_cast<char *>(p);
      slab_left_ = slab;
    }
    void *p = slab_cur_;
    slab_cur_ += c;
    slab_left_ -= c;
    return p;
  }
  void release(void *p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu_);
    free_[cap].push_back(p);
  }
 private:
  static constexpr size_t kSlab = 64u << 20, kMaxPinned = 8ull << 30;
  std::mutex mu_;
  std::map<size_t, std::vector<void *>> free_;
  char *slab_cur_ = nullptr;
  size_t slab_left_ = 0, pinned_total_ = 0;
};

// accept_waveform is a host memcpy of the caller's samples into pinned memory; one core moves ~8 GB/s, so a 256-stream
// batch (180 MB) spends >20 ms there. Large copies are split over a few helper threads that keep spinning for a short
// while after a job (an accept loop issues the next one within microseconds) and otherwise sleep on a condition variable.
// Helper threads a rank may start for a short burst of independent host work (beside the calling thread): half of its share of
// the cores (affinity- and torchrun-aware), at most 7, none when the share is below 4.
static int host_helper_threads() {
  static const int n = [] {
    int hw = (int)std::thread::hardware_concurrency();
    cpu_set_t cs;
    if (sched_getaffinity(0, sizeof cs, &cs) == 0 && CPU_COUNT(&cs) > 0) hw = std::min(hw > 0 ? hw : CPU_COUNT(&cs), CPU_COUNT(&cs));
    int ranks = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    const int share = std::max(1, hw / ranks);
    if (const char *e = getenv("B200ASR_PLAN_THREADS")) return std::max(0, atoi(e) - 1);
    return share >= 4 ? std::min(7, share / 2 - 1) : 0;
  }();
  return n;
}

class ParallelCopy {
 public:
  static ParallelCopy &get() { static ParallelCopy p; return p; }
  void copy(void *dst, const void *src, size_t bytes) {
    if (n_workers_ == 0 || bytes < kMinBytes) { memcpy(dst, src, bytes); return; }
    std::lock_guard<std::mutex> serial(call_mu_);       // one job at a time
    const int parts = n_workers_ + 1;
    const size_t chunk = ((bytes / parts) + 4095) & ~size_t(4095);
    dst_ = static_cast<char *>(dst); src_ = static_cast<const char *>(src); bytes_ = bytes; chunk_ = chunk;
    pending_.store(n_workers_, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    run_part(0);
    while (pending_.load(std::memory_order_acquire) != 0) { /* spin: the parts are equal */ }
  }
  ~ParallelCopy() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; epoch_.fetch_add(1, std::memory_order_release); }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
  }
 private:
  static constexpr size_t kMinBytes = 256 * 1024;
  ParallelCopy() {
    // cores this process may use (affinity / cgroup aware), shared with the other ranks of a torchrun job on the box
    int hw = (int)std::thread::hardware_concurrency();
    cpu_set_t cs;
    if (sched_getaffinity(0, sizeof cs, &cs) == 0 && CPU_COUNT(&cs) > 0) hw = std::min(hw > 0 ? hw : CPU_COUNT(&cs), CPU_COUNT(&cs));
    int ranks = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    const int share = std::max(1, hw / ranks);
    // helpers besides the calling thread. A rank's kernel-launch thread must never wait for a core (the search issues a launch
    // every ~11 us), and a feeder thread usually runs accept_waveform beside it; so helpers take at most half of the rank's share
    // of the cores minus those two, and none at all when the share is small (8 ranks on 32 cores: measured, 40 runnable threads
    // on 32 cores stretched the device-timed step by 7 % and the accept loop fourfold; one copying thread per rank hides
    // under the decode of the previous batch anyway)
    int n = share >= 8 ? std::min(7, share / 2 - 1) : 0;
    if (const char *e = getenv("B200ASR_COPY_THREADS")) n = atoi(e) - 1;
    n_workers_ = std::max(0, std::min(n, share > 1 ? share - 1 : 0));
    for (int i = 0; i < n_workers_; ++i) threads_.emplace_back([this, i] { worker(i + 1); });
  }
  void run_part(int part) {
    const size_t b = (size_t)part * chunk_;
    if (b < bytes_) memcpy(dst_ + b, src_ + b, std::min(chunk_, bytes_ - b));
  }
  void worker(int part) {
    unsigned long long seen = 0;
    for (;;) {
      // spin briefly for the next job, then block
      const auto t0 = std::chrono::steady_clock::now();
      while (epoch_.load(std::memory_order_acquire) == seen) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
          break;
        }
      }
      seen = epoch_.load(std::memory_order_acquire);
      if (stop_) return;
      run_part(part);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> threads_;
  int n_workers_ = 0;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_;
  std::atomic<unsigned long long> epoch_{0};
  std::atomic<int> pending_{0};
  bool stop_ = false;
  char *dst_ = nullptr; const char *src_ = nullptr; size_t bytes_ = 0, chunk_ = 0;
};

// Device blocks for the accept-time PCM upload, one pool and one copy stream per device: accept_waveform enqueues the
// host->device copy of the samples it has just taken, so by the time decode_streams is called the PCM of the batch is
// (mostly) resident and the pass starts with fbank instead of a 180 MB transfer. Blocks are power-of-two size classes
// carved from 256 MB slabs and recycled; one copy stream per device keeps reuse of a recycled block stream-ordered.
class DevicePcmPool {
 public:
  static DevicePcmPool &get(int device) {
    static std::mutex m;
    static std::map<int, std::unique_ptr<DevicePcmPool>> pools;
    std::lock_guard<std::mutex> lk(m);
    auto &p = pools[device];
    if (!p) p.reset(new DevicePcmPool(device));
    return *p;
  }
  cudaStream_t stream() {
    std::lock_guard<std::mutex> lk(mu_);
    if (!st_ && cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); st_ = nullptr; }
    return st_;
  }
  float *alloc(size_t bytes, size_t *cap) {
    size_t c = 64 * 1024;
    while (c < bytes) c <<= 1;
    *cap = c;
    std::lock_guard<std::mutex> lk(mu_);
    auto &fl = free_[c];
    if (!fl.empty()) { float *p = fl.back(); fl.pop_back(); return p; }
    if (c > slab_left_) {
      const size_t slab = std::max(c, kSlab);
      void *p = nullptr;
      if (total_ + slab > kMax || cudaMalloc(&p, slab) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      total_ += slab;
      slab_cur_ = reinterpret_cast<char *>(p);
      slab_left_ = slab;
    }
    float *p = reinterpret_cast<float *>(slab_cur_);
    slab_cur_ += c;
    slab_left_ -= c;
    return p;
  }
  void release(float *p, size_t cap) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu_);
    free_[cap].push_back(p);
  }
 private:
  explicit DevicePcmPool(int) {}
  static constexpr size_t kSlab = 256u << 20, kMax = 16ull << 30;
  std::mutex mu_;
  std::map<size_t, std::vector<float *>> free_;
  char *slab_cur_ = nullptr;
  size_t slab_left_ = 0, total_ = 0;
  cudaStream_t st_ = nullptr;
};

struct PinnedSamples {
  float *p = nullptr;
  size_t n = 0, cap_bytes = 0;
  // device copy maintained by accept_waveform (null when the pool is exhausted or uploads are disabled)
  float *d = nullptr;
  size_t dcap_bytes = 0, n_up = 0;
  int device = -1;
  bool in_flight = false;   // uploads queued since the last pass that waited for them (decode clears it)
  // a pinned block must not go back to the pool (and be overwritten by another stream) while a DMA still reads it
  void settle() {
    if (in_flight && device >= 0) {
      if (cudaSetDevice(device) == cudaSuccess) cudaStreamSynchronize(DevicePcmPool::get(device).stream());
      cudaGetLastError();
      in_flight = false;
    }
  }
  ~PinnedSamples() {
    settle();
    PinnedPool::get().release(p, cap_bytes);
    if (d) DevicePcmPool::get(device).release(d, dcap_bytes);
  }
  bool on_device() const { return d != nullptr && n_up == n; }
  // enqueue the upload of the samples not yet on the device (copy stream of `dev`)
  void upload(int dev) {
    static const bool enabled = getenv("B200ASR_NO_EAGER_UPLOAD") == nullptr;
    if (!enabled || n == 0) return;
    DevicePcmPool &pool = DevicePcmPool::get(dev);
    cudaStream_t cst = pool.stream();
    if (!cst) return;
    if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return; }
    if (n * sizeof(float) > dcap_bytes) {
      size_t ncap = 0;
      float *nd = pool.alloc(std::max(cap_bytes, n * sizeof(float)), &ncap);
      if (d) pool.release(d, dcap_bytes);
      d = nd; dcap_bytes = nd ? ncap : 0; n_up = 0; device = dev;
      if (!nd) return;
    }
    if (cudaMemcpyAsync(d + n_up, p + n_up, (n - n_up) * sizeof(float), cudaMemcpyHostToDevice, cst) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    n_up = n;
    in_flight = true;
  }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  const float *data() const { return p; }
  void append(const float *src, size_t cnt) {
    if ((n + cnt) * sizeof(float) > cap_bytes) {
      size_t ncap = 0;
      float *np_ = reinterpret_cast<float *>(PinnedPool::get().alloc((n + cnt) * sizeof(float) * (p ? 2 : 1), &ncap));
      if (n) memcpy(np_, p, n * sizeof(float));
      settle();
      PinnedPool::get().release(p, cap_bytes);
      p = np_;
      cap_bytes = ncap;
    }
    ParallelCopy::get().copy(p + n, src, cnt * sizeof(float));
    n += cnt;
  }
};

// A hotword automaton with its device copy (the recognizer's own, or one attached to a stream by
// B200AsrCreateOfflineStreamWithHotwords)
struct GraphPair {
  ContextGraphHost host;
  ContextGraphDev dev;
  int device = 0;
  bool on_device = false;
  void upload(int dev_id);
  ~GraphPair() {
    if (!on_device) return;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
    int *ptrs[] = {dev.edge_start, dev.edge_token, dev.edge_child, dev.fail, dev.token, dev.is_end, dev.output};
    for (int *p : ptrs) if (p) cudaFree(p);
    double *dp[] = {dev.token_score, dev.node_score, dev.output_score};
    for (double *p : dp) if (p) cudaFree(p);
  }
};
void GraphPair::upload(int dev_id) {
  device = dev_id;
  const int N = host.n_nodes();
  if (N <= 1) return;
  auto up_i = [&](const std::vector<int> &v) { int *d; CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(int)));
    CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice)); return d; };
  auto up_d = [&](const std::vector<double> &v) { double *d; CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(double)));
    CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice)); return d; };
  on_device = true;
  dev.n_nodes = N;
  dev.edge_start = up_i(host.edge_start); dev.edge_token = up_i(host.edge_token);
  dev.edge_child = up_i(host.edge_child); dev.fail = up_i(host.fail); dev.token = up_i(host.token);
  dev.is_end = up_i(host.is_end); dev.output = up_i(host.output);
  dev.token_score = up_d(host.token_score); dev.node_score = up_d(host.node_score);
  dev.output_score = up_d(host.output_score);
}

struct Engine {
  // config
  std::map<std::string, std::string> cfg;
  std::vector<int> num_layers, ds_factor, enc_dim, ff_dim, num_heads, cnn_kernel;
  int qd = 32, pd = 4, vd = 12, pos_dim = 48, feat_dim = 80, dec_dim = 512, join_dim = 512, ctx_size = 2, V = 2000;
  int blank_id = 0, unk_id = 2, out_dim = 512;
  int device = 0, precision = 0;
  std::string decoding_method = "greedy_search";
  int max_active_paths = 4;
  float hotwords_score = 1.5f, blank_penalty = 0.f;
  std::vector<std::string> id2token;

  std::map<std::string, Tensor> tensors;
  std::vector<StackW> stacks;
  // embed weights (re-laid out)
  float *w_conv0 = nullptr, *w_conv1 = nullptr, *w_conv2 = nullptr, *w_dw7 = nullptr, *w_out = nullptr;
  std::vector<float *> owned;   // extra device allocations to free
  std::map<const float *, const float *> w_lo;   // weight -> its pre-split low part (3xTF32 mode)
  struct W16 { void *hi, *lo; int ld; };
  std::map<const float *, W16> w16;              // weight -> its 16-bit operand copies (fp16 split of the FP32 mode, or bf16)
  bool use_f16x3 = false;                        // FP32 mode on the fp16 operand split instead of 3xTF32
  typedef void (*GemmFn)(const GemmArgs &, cudaStream_t);
  GemmFn gemm_fn() const { return precision == 1 ? launch_gemm_tc : precision == 0 ? launch_gemm_tc3 : precision == 3 ? launch_gemm_bf16 : launch_gemm_fp32; }
  const W16 &w16_for(const float *Wt, int N, int K) {
    auto it = w16.find(Wt);
    if (it == w16.end()) {
      W16 w{nullptr, nullptr, 0};
      split_weights_16(Wt, N, K, precision == 3, &w.hi, &w.lo, &w.ld, st);
      it = w16.emplace(Wt, w).first;
    }
    return it->second;
  }

  FbankTables fb{};
  cudaStream_t st = nullptr;
  cudaEvent_t ev[8]{};
  cudaEvent_t ev_copy = nullptr;   // marks the accept-time uploads a pass has to wait for
  SearchState *search = nullptr;   // lane 0's (also used by the raw B200AsrBeamSearch entry point)
  Lane lanes[kMaxLanes];
  // Spatial split of the SMs for pipelined passes (CUDA green contexts): `search_sms` SMs run the search lanes, the rest
  // run the encoder of the following groups. A persistent encoder CTA holds its SM for a whole kernel, so on shared SMs the
  // search's small kernels and the encoder's tiles delay each other; on disjoint SMs neither sees the other.
  struct SmPartition {
    CUgreenCtx g_search = nullptr, g_enc = nullptr;
    int search_sms = 0, enc_sms = 0;
    bool tried = false, ok = false;
  } part;
  cudaStream_t st_part = nullptr;     // encoder stream inside the encoder partition
  cudaEvent_t ev_part = nullptr;
  void ensure_partition();
  int n_groups_last = 1;
  // Pinned staging for host->device descriptor uploads larger than the driver's inline limit (the per-stack tensor maps, 160 KB
  // for 256 utterances): from PAGEABLE memory such a cudaMemcpyAsync first waits for the stream to drain, which tied the host to
  // the GPU at every stack boundary (26 ms of "host time" per encoder) and, in chained passes, held back the launch sequence of
  // the previous batch's search until the next encoder had all but finished. Bump-allocated, reset when a pass starts.
  unsigned char *pin_arena = nullptr;
  size_t pin_cap = 0, pin_used = 0;
  unsigned char *pinned_scratch(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (!pin_arena) {
      pin_cap = (size_t)32 << 20;
      if (cudaHostAlloc(reinterpret_cast<void **>(&pin_arena), pin_cap, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); pin_arena = nullptr; pin_cap = 0; }
    }
    if (!pin_arena || pin_used + bytes > pin_cap) return nullptr;     // caller falls back to pageable memory
    unsigned char *p = pin_arena + pin_used;
    pin_used += bytes;
    return p;
  }
  double hp_gemm = 0, hp_plan = 0, hp_attn = 0;   // B200ASR_HOST_PROF: host time inside gemm(), build_attn_plan(), attention launches
  float search_busy_ms = 0, lane_ms[kMaxLanes] = {0};
  float lane_t[kMaxLanes][4] = {{0}};   // last pass, ms from its start: encoder begin / end, search begin / end of each group
  std::vector<int> lane_of, idx_of;   // last pass: utterance -> (lane, index inside the lane's group)
  double kappa = 50.0;                // search ms per second of utterance length / encoder ms per audio-second (adapted per pass)
  long long d2h_bytes_last = 0;
  SearchModel sm{};
  // Decoder outputs of all V^2 two-token contexts (the reference's dec_cache, core/asr_engine.py:1072-1088, filled completely
  // and ahead of time): built on the first search when it fits comfortably, see ensure_dec_table()
  bool dec_table_tried = false;
  void ensure_dec_table();
  // Single-pass attention weights (attn_weights_tc.cu ONEPASS): unnormalised weights + row sums, the consumers divide. A row
  // sum outside the safe range raises sm_flag on the device; the pass is then repeated with the exact two-pass kernel, which
  // the recognizer keeps from there on (B200ASR_SOFTMAX_2PASS=1 selects it from the start).
  bool exact_softmax = false;
  int *sm_flag = nullptr;        // device
  int *sm_flag_host = nullptr;   // pinned mirror, copied at the end of every encoder run
  DevBuf b_ls;
  bool softmax_overflowed();
  std::shared_ptr<GraphPair> graph;          // the recognizer's hotword automaton (null = none)
  const ContextGraphDev *pass_graph = nullptr;   // what the pass being issued scores with (decode sets it per partition)
  const ContextGraphHost &cg_host() const { static const ContextGraphHost empty; return graph ? graph->host : empty; }
  std::mutex mu;
  Timings tm;
  long long launches_last = 0;
  bool profiling = false;
  double gemm_ms = 0, gemm_flops = 0, gemm_bytes = 0;
  long long gemm_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
  size_t gemm_ev_used = 0;

  // workspaces
  DevBuf b_pcm, b_soff, b_foff, b_feats, b_featin;
  // one tile counter per GEMM launch of a pass (dynamic tile scheduling of the persistent kernels), zeroed at pass start
  static constexpr int kTileCounters = 16384;
  DevBuf b_counters;
  int counter_next = 0;
  int *next_tile_counter() {
    static const bool dyn_tiles = getenv("B200ASR_STATIC_TILES") == nullptr;
    return (dyn_tiles && b_counters.p && counter_next < kTileCounters) ? b_counters.ptr<int>() + counter_next++ : nullptr;
  }
  void reset_tile_counters() {
    int *c = b_counters.get<int>(kTileCounters);
    CUDA_CHECK(cudaMemsetAsync(c, 0, kTileCounters * sizeof(int), st));
    counter_next = 0;
  }
  DevBuf b_T, b_c0off, b_c1off, b_len[4], b_off[4], b_aoff;
  DevBuf b_c0, b_c1, b_c2, b_dw, b_pw1, b_cn, b_x0, b_xc, b_sin, b_w1, b_proj, b_hid, b_A, b_pe, b_pp, b_cat, b_enc;
  DevBuf b_stack[6];
  DevBuf b_tmp, b_tmp2;
  // host-side batch description of the last encoder run
  std::vector<int> h_T, h_T1, h_Tp;
  std::vector<std::vector<int>> h_len = std::vector<std::vector<int>>(4), h_off = std::vector<std::vector<int>>(4);
  int last_n = 0;
  // staged batches for device-resident benchmarking
  struct Staged { float *pcm = nullptr; long long *soff = nullptr; std::vector<long long> h_soff; int n = 0; };
  std::map<int, Staged> staged;
  int next_handle = 0;

  // Host staging data of the asynchronous descriptor uploads of one pass: kept alive until the next pass starts, so
  // no upload needs a stream synchronisation (each one used to stall the GPU while the host prepared the next plan).
  std::vector<std::shared_ptr<void>> host_keep;
  template <typename T>
  const T *keep(std::vector<T> &&v) {
    auto p = std::make_shared<std::vector<T>>(std::move(v));
    host_keep.push_back(p);
    return p->data();
  }
  ~Engine();
  void load(const B200AsrOfflineRecognizerConfig *c);
  void load_container(const std::string &path, const std::string &prefix);
  const float *W(const std::string &name, std::vector<int> expect = {});
  float *upload(const std::vector<float> &h);
  void gemm(const float *A, int lda, const float *Wt, const float *bias, const float *R, int ldr, float *C, int ldc, int M, int N,
            int K, int act);
  void feed_forward(const LayerW &w, int j, const float *x, const float *R, float *out, int M, int D, float *hidden);
  void set_graph(const int32_t *tokens, const int32_t *offsets, const float *scores, int n);
  // pipeline pieces (device pointers)
  // h_len[u] = samples of utterance u; d_slen null = packed PCM (lengths from consecutive offsets)
  void run_fbank(const float *d_pcm, const long long *d_soff, const long long *d_slen, const std::vector<long long> &h_len, int n,
                 float **d_feats, std::vector<int> *T);
  void run_encoder(const float *d_feats, const std::vector<int> &T, DevBuf &enc_buf, float **d_enc, std::vector<int> *Tp);
  Lane &lane(int i);
  std::vector<std::vector<int>> plan_groups(const std::vector<long long> &h_len) const;
  struct AttnPlan {   // tensor-core attention application: per-stack maps / offsets (valid for every layer of the stack)
    bool use = false;
    const void *mapsA = nullptr, *mapsV12 = nullptr, *mapsV12lo = nullptr, *mapsVh = nullptr, *mapsVhlo = nullptr;
    const int *tile_off12 = nullptr, *tile_offh = nullptr;
    const long long *vt_off12 = nullptr, *vt_offh = nullptr;
    float *VT12 = nullptr, *VT12lo = nullptr, *VTh = nullptr, *VThlo = nullptr;
    int n_tiles12 = 0, n_tilesh = 0;
    const int *dw_tile_off = nullptr;   // cumulative ceil(len / 128) per utterance at the stack's rate (conv module tiles)
    int dw_tiles = 0;
  };
  DevBuf b_up[4], b_down[4], b_dwt;
  DevBuf b_maps, b_tileoff, b_vtoff, b_vt12, b_vt12lo, b_vth, b_vthlo;
  void build_attn_plan(AttnPlan *pl, const StackW &s, int q, int n, const std::vector<long long> &aoff_host);
  void attn_apply(const AttnPlan &pl, const StackW &s, const RaggedDesc &r, const long long *aoff, const float *X, int ldx,
                  const float *S, int lds, const float *Y, int ldy, int C, bool single_head, float *out, int ldo, const float *Ls = nullptr);
  void run_layer(const StackW &s, const LayerW &w, float *src, const RaggedDesc &r, const long long *aoff, int M, int Lmax,
                 const AttnPlan &pl);
  void decode(Stream *const *ss, int n);
  void decode_part(Stream *const *ss, int n);
  void unpack_results(Stream *const *ss, int nb, const std::vector<int> &Tp);
  const ContextGraphDev *default_graph() const { return (graph && graph->on_device) ? &graph->dev : nullptr; }
  std::shared_ptr<GraphPair> make_graph(const int32_t *tokens, const int32_t *offsets, const float *scores, int n);
  // h_soff[u] = first sample of utterance u relative to d_pcm, h_len[u] = its samples. Results stay in the lanes' search
  // states (lane_of / idx_of map utterances to them) until the next pass.
  void decode_pcm_device(const float *d_pcm, const std::vector<long long> &h_soff, const std::vector<long long> &h_len, int n,
                         std::vector<int> *Tp, const float *d_featin = nullptr, const std::vector<long long> *h_foff = nullptr,
                         const std::vector<std::vector<int>> *forced_groups = nullptr);
  struct UttResult { int n_tokens; const int *tokens, *frames; const float *tok_lp, *stats; };
  UttResult result_of(int u) const;
  void collect_gemm_times();
};

struct Stream {
  Engine *eng;
  PinnedSamples samples;
  std::shared_ptr<GraphPair> graph;   // per-stream hotwords (null = the recognizer's)
  // precomputed features in place of samples (B200AsrAcceptFeaturesOffline: ROVER's shared fbank, core/asr_engine.py:2346-2350)
  std::vector<float> feats;
  int feat_T = 0;
  long long feat_samples = 0;
  bool has_feats() const { return feat_T > 0; }
  long long n_samples() const { return has_feats() ? feat_samples : (long long)samples.size(); }
  // result storage
  B200AsrOfflineRecognizerResult res{};
  std::string text, json;
  std::vector<std::string> tok_str;
  std::vector<const char *> tok_ptr;
  std::vector<int32_t> token_ids, frames;
  std::vector<float> timestamps, lps, tsallis, margin, entropy, top1;
  bool decoded = false;
  bool json_built = false;
};

Engine::~Engine() {
  if (pin_arena) cudaFreeHost(pin_arena);
  if (sm_flag) cudaFree(sm_flag);
  if (sm_flag_host) cudaFreeHost(sm_flag_host);
  for (auto &kv : tensors) if (kv.second.dev) cudaFree(kv.second.dev);
  for (float *p : owned) cudaFree(p);
  for (auto &kv : w16) { cudaFree(kv.second.hi); cudaFree(kv.second.lo); }
  if (fb.window) fbank_tables_destroy(&fb);
  for (int i = 0; i < kMaxLanes; ++i) {
    Lane &l = lanes[i];
    if (l.search && l.search != search) search_state_destroy(l.search);
    cudaEvent_t evs[] = {l.f0, l.f1, l.e1, l.s0, l.s1};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (l.st) cudaStreamDestroy(l.st);
  }
  if (search) search_state_destroy(search);
  for (auto &e : ev) if (e) cudaEventDestroy(e);
  if (ev_copy) cudaEventDestroy(ev_copy);
  for (auto &p : gemm_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  for (auto &kv : staged) { cudaFree(kv.second.pcm); cudaFree(kv.second.soff); }
  graph.reset();
  if (st_part) cudaStreamDestroy(st_part);
  if (ev_part) cudaEventDestroy(ev_part);
  if (part.g_search) green_api().GreenCtxDestroy(part.g_search);
  if (part.g_enc) green_api().GreenCtxDestroy(part.g_enc);
  if (st) cudaStreamDestroy(st);
}

void Engine::load_container(const std::string &path, const std::string &prefix) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open weight container: " + path);
  std::string first;
  std::getline(f, first);
  char magic[32];
  int ver = 0;
  long long hb = 0;
  if (sscanf(first.c_str(), "%31s %d %lld", magic, &ver, &hb) != 3 || std::string(magic) != "B200ASRW")
    throw std::runtime_error(path + ": not a B200ASRW container");
  f.seekg(0);
  std::string header((size_t)hb, '\0');
  f.read(&header[0], hb);
  std::stringstream hs(header);
  std::string line;
  std::getline(hs, line);
  struct Ent { std::string name; std::vector<int> shape; long long off, nbytes; };
  std::vector<Ent> ents;
  while (std::getline(hs, line)) {
    std::stringstream ls(line);
    std::string kind;
    ls >> kind;
    if (kind == "end" || kind.empty() || kind[0] == '\0') break;
    if (kind == "config") {
      std::string k, v;
      ls >> k >> v;
      if (!cfg.count(k)) cfg[k] = v;
      else if (cfg[k] != v && k != "name") throw std::runtime_error("containers disagree on config " + k);
    } else if (kind == "tensor") {
      Ent e;
      std::string dt;
      int nd;
      ls >> e.name >> dt >> nd;
      e.shape.resize(nd);
      for (int i = 0; i < nd; ++i) ls >> e.shape[i];
      ls >> e.off >> e.nbytes;
      if (dt != "f32") throw std::runtime_error("unsupported dtype in container: " + dt);
      if (e.name.compare(0, prefix.size(), prefix) == 0) ents.push_back(e);
    }
  }
  for (auto &e : ents) {
    Tensor t;
    t.shape = e.shape;
    t.host.resize((size_t)e.nbytes / 4);
    f.seekg(hb + e.off);
    f.read(reinterpret_cast<char *>(t.host.data()), e.nbytes);
    if (!f) throw std::runtime_error(path + ": truncated tensor " + e.name);
    if (t.numel() * 4 != (size_t)e.nbytes) throw std::runtime_error(path + ": shape/bytes mismatch for " + e.name);
    CUDA_CHECK(cudaMalloc(&t.dev, std::max<size_t>(e.nbytes, 16)));
    CUDA_CHECK(cudaMemcpy(t.dev, t.host.data(), e.nbytes, cudaMemcpyHostToDevice));
    tensors[e.name] = std::move(t);
  }
  if (ents.empty()) throw std::runtime_error(path + ": holds no '" + prefix + "*' tensors");
}

const float *Engine::W(const std::string &name, std::vector<int> expect) {
  auto it = tensors.find(name);
  if (it == tensors.end()) throw std::runtime_error("missing tensor: " + name);
  if (!expect.empty() && it->second.shape != expect) {
    std::string s = "tensor " + name + " has shape [";
    for (int d : it->second.shape) s += std::to_string(d) + ",";
    s += "] expected [";
    for (int d : expect) s += std::to_string(d) + ",";
    throw std::runtime_error(s + "]");
  }
  return it->second.dev;
}

float *Engine::upload(const std::vector<float> &h) {
  float *d;
  CUDA_CHECK(cudaMalloc(&d, h.size() * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  owned.push_back(d);
  return d;
}

void Engine::load(const B200AsrOfflineRecognizerConfig *c) {
  const auto &mc = c->model_config;
  auto str = [](const char *s) { return std::string(s ? s : ""); };
  const std::string prov = str(mc.provider);
  if (!prov.empty() && prov != "cuda" && prov != "b200")
    throw std::runtime_error("provider '" + prov + "' is not available: this library has only the CUDA (sm_100a) path");
  if (c->feat_config.sample_rate != 0 && c->feat_config.sample_rate != 16000)
    throw std::runtime_error("only 16 kHz input is supported (core/asr_engine.py:706)");
  if (c->feat_config.feature_dim != 0 && c->feat_config.feature_dim != 80)
    throw std::runtime_error("only 80-bin fbank is supported (core/asr_engine.py:710)");
  device = c->device_id;
  precision = c->precision;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    throw std::runtime_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libb200asr has no CPU fallback)");
  if (device < 0 || device >= ndev) throw std::runtime_error("device_id out of range");
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) throw std::runtime_error("libb200asr is built for sm_100a only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
  CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CUDA_CHECK(cudaMalloc(&sm_flag, sizeof(int)));
  CUDA_CHECK(cudaMemset(sm_flag, 0, sizeof(int)));
  CUDA_CHECK(cudaHostAlloc(&sm_flag_host, sizeof(int), cudaHostAllocPortable));
  *sm_flag_host = 0;
  exact_softmax = getenv("B200ASR_SOFTMAX_2PASS") != nullptr && atoi(getenv("B200ASR_SOFTMAX_2PASS")) != 0;
  for (auto &x : ev) CUDA_CHECK(cudaEventCreate(&x));
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_copy, cudaEventDisableTiming));

  load_container(str(mc.transducer.encoder), "encoder.");
  load_container(str(mc.transducer.decoder), "decoder.");
  load_container(str(mc.transducer.joiner), "joiner.");
  auto geti = [&](const char *k) { if (!cfg.count(k)) throw std::runtime_error(std::string("container lacks config ") + k); return atoi(cfg[k].c_str()); };
  num_layers = parse_int_list(cfg["num_encoder_layers"]);
  ds_factor = parse_int_list(cfg["downsampling_factor"]);
  enc_dim = parse_int_list(cfg["encoder_dim"]);
  ff_dim = parse_int_list(cfg["feedforward_dim"]);
  num_heads = parse_int_list(cfg["num_heads"]);
  cnn_kernel = parse_int_list(cfg["cnn_module_kernel"]);
  qd = geti("query_head_dim"); pd = geti("pos_head_dim"); vd = geti("value_head_dim"); pos_dim = geti("pos_dim");
  feat_dim = geti("feature_dim"); dec_dim = geti("decoder_dim"); join_dim = geti("joiner_dim"); ctx_size = geti("context_size");
  V = geti("vocab_size"); blank_id = geti("blank_id"); unk_id = geti("unk_id");
  const size_t ns = num_layers.size();
  if (ns == 0 || ns > 6 || ds_factor.size() != ns || enc_dim.size() != ns || ff_dim.size() != ns || num_heads.size() != ns ||
      cnn_kernel.size() != ns)
    throw std::runtime_error("inconsistent stack configuration in container");
  if (ctx_size != 2) throw std::runtime_error("only context_size=2 is built");
  if ((dec_dim & 3) || (join_dim & 3)) throw std::runtime_error("decoder_dim and joiner_dim must be multiples of 4");
  if (feat_dim != 80) throw std::runtime_error("only feature_dim=80 is built");
  out_dim = *std::max_element(enc_dim.begin(), enc_dim.end());

  // ---- embed weights, re-laid out channels-last
  {
    const auto &w0 = tensors.at("encoder.embed.conv0.weight").host;   // [8,1,3,3] -> [3][3][8]
    W("encoder.embed.conv0.weight", {8, 1, 3, 3});
    std::vector<float> t(72);
    for (int co = 0; co < 8; ++co) for (int k = 0; k < 9; ++k) t[k * 8 + co] = w0[co * 9 + k];
    w_conv0 = upload(t);
    const auto &w1 = tensors.at("encoder.embed.conv1.weight").host;   // [32,8,3,3] -> [3][3][8][32]
    W("encoder.embed.conv1.weight", {32, 8, 3, 3});
    t.assign(72 * 32, 0.f);
    for (int co = 0; co < 32; ++co) for (int ci = 0; ci < 8; ++ci) for (int k = 0; k < 9; ++k)
      t[(k * 8 + ci) * 32 + co] = w1[(co * 8 + ci) * 9 + k];
    w_conv1 = upload(t);
    const auto &w2 = tensors.at("encoder.embed.conv2.weight").host;   // [128,32,3,3] -> GEMM weight [128][(kh,kw,ci)]
    W("encoder.embed.conv2.weight", {128, 32, 3, 3});
    t.assign(288 * 128, 0.f);
    for (int co = 0; co < 128; ++co) for (int ci = 0; ci < 32; ++ci) for (int k = 0; k < 9; ++k)
      t[(size_t)co * 288 + k * 32 + ci] = w2[(co * 32 + ci) * 9 + k];
    w_conv2 = upload(t);
    const auto &wd = tensors.at("encoder.embed.convnext.dw.weight").host;  // [128,1,7,7] -> [7][7][128]
    W("encoder.embed.convnext.dw.weight", {128, 1, 7, 7});
    t.assign(49 * 128, 0.f);
    for (int c2 = 0; c2 < 128; ++c2) for (int k = 0; k < 49; ++k) t[k * 128 + c2] = wd[c2 * 49 + k];
    w_dw7 = upload(t);
    const int D0 = enc_dim[0];
    const auto &wo = tensors.at("encoder.embed.out.weight").host;     // [D0, c*19+f] -> [D0, f*128+c]
    W("encoder.embed.out.weight", {D0, 128 * 19});
    t.assign((size_t)D0 * 2432, 0.f);
    for (int o = 0; o < D0; ++o) for (int c2 = 0; c2 < 128; ++c2) for (int f = 0; f < 19; ++f)
      t[(size_t)o * 2432 + f * 128 + c2] = wo[(size_t)o * 2432 + c2 * 19 + f];
    w_out = upload(t);
  }
  // ---- stacks
  stacks.resize(ns);
  for (size_t i = 0; i < ns; ++i) {
    StackW &s = stacks[i];
    s.L = num_layers[i]; s.ds = ds_factor[i]; s.D = enc_dim[i]; s.F = ff_dim[i]; s.H = num_heads[i]; s.k = cnn_kernel[i];
    const std::string sp = "encoder.stack" + std::to_string(i) + ".";
    if (s.ds > 1) {
      s.ds_bias = W(sp + "downsample.bias", {s.ds});
      s.combiner = W(sp + "out_combiner.scale", {s.D});
    }
    const int D = s.D, H = s.H, h = (3 * D) / 4;
    for (int l = 0; l < s.L; ++l) {
      const std::string p = sp + "layer" + std::to_string(l) + ".";
      LayerW w{};
      w.attn_in_w = W(p + "attn_w.in_proj.weight", {H * (2 * qd + pd), D});
      w.attn_in_b = W(p + "attn_w.in_proj.bias", {H * (2 * qd + pd)});
      w.pos_w = W(p + "attn_w.linear_pos.weight", {H * pd, pos_dim});
      const int fd[3] = {(s.F * 3) / 4, s.F, (s.F * 5) / 4};
      for (int j = 0; j < 3; ++j) {
        const std::string q = p + "ff" + std::to_string(j + 1) + ".";
        w.ff_dim[j] = fd[j];
        w.ff_in_w[j] = W(q + "in.weight", {fd[j], D}); w.ff_in_b[j] = W(q + "in.bias", {fd[j]});
        w.ff_out_w[j] = W(q + "out.weight", {D, fd[j]}); w.ff_out_b[j] = W(q + "out.bias", {D});
      }
      w.nl_in_w = W(p + "nonlin.in.weight", {3 * h, D}); w.nl_in_b = W(p + "nonlin.in.bias", {3 * h});
      w.nl_out_w = W(p + "nonlin.out.weight", {D, h}); w.nl_out_b = W(p + "nonlin.out.bias", {D});
      for (int j = 0; j < 2; ++j) {
        const std::string a = p + "attn" + std::to_string(j + 1) + ".";
        w.sa_in_w[j] = W(a + "in.weight", {H * vd, D}); w.sa_in_b[j] = W(a + "in.bias", {H * vd});
        w.sa_out_w[j] = W(a + "out.weight", {D, H * vd}); w.sa_out_b[j] = W(a + "out.bias", {D});
        const std::string cname = p + "conv" + std::to_string(j + 1) + ".";
        w.cv_in_w[j] = W(cname + "in.weight", {2 * D, D}); w.cv_in_b[j] = W(cname + "in.bias", {2 * D});
        W(cname + "dw.weight", {D, 1, s.k});
        const auto &dw = tensors.at(cname + "dw.weight").host;       // [D,1,k] -> [k][D]
        std::vector<float> t((size_t)s.k * D);
        for (int c2 = 0; c2 < D; ++c2) for (int j2 = 0; j2 < s.k; ++j2) t[(size_t)j2 * D + c2] = dw[(size_t)c2 * s.k + j2];
        w.cv_dw_w[j] = upload(t);
        w.cv_dw_b[j] = W(cname + "dw.bias", {D});
        w.cv_out_w[j] = W(cname + "out.weight", {D, D}); w.cv_out_b[j] = W(cname + "out.bias", {D});
      }
      w.norm_bias = W(p + "norm.bias", {D}); w.norm_log_scale = W(p + "norm.log_scale", {1});
      w.bypass = W(p + "bypass.scale", {D}); w.bypass_mid = W(p + "bypass_mid.scale", {D});
      s.layers.push_back(w);
    }
  }
  W("encoder.downsample_output.bias", {2});
  W("encoder.encoder_proj.weight", {join_dim, out_dim});
  // ---- decoder / joiner
  sm.emb = W("decoder.embedding.weight", {V, dec_dim});
  sm.conv_w = W("decoder.conv.weight", {dec_dim, 4, ctx_size});
  sm.dec_proj_w = W("decoder.decoder_proj.weight", {join_dim, dec_dim});
  sm.dec_proj_b = W("decoder.decoder_proj.bias", {join_dim});
  sm.join_w = W("joiner.output_linear.weight", {V, join_dim});
  sm.join_b = W("joiner.output_linear.bias", {V});
  sm.V = V; sm.dd = dec_dim; sm.jd = join_dim; sm.blank_id = blank_id; sm.unk_id = unk_id;
  {
    // per-token halves of the decoder's grouped convolution (groups of 4 channels, kernel 2)
    const auto &emb = tensors.at("decoder.embedding.weight").host;
    const auto &cw = tensors.at("decoder.conv.weight").host;   // [dd, 4, 2]
    std::vector<float> p0((size_t)V * dec_dim), p1((size_t)V * dec_dim);
    for (int y = 0; y < V; ++y)
      for (int o = 0; o < dec_dim; ++o) {
        const int g4 = (o >> 2) << 2;
        float a0 = 0.f, a1 = 0.f;
        for (int i = 0; i < 4; ++i) {
          a0 = fmaf(cw[(size_t)o * 8 + i * 2 + 0], emb[(size_t)y * dec_dim + g4 + i], a0);
          a1 = fmaf(cw[(size_t)o * 8 + i * 2 + 1], emb[(size_t)y * dec_dim + g4 + i], a1);
        }
        p0[(size_t)y * dec_dim + o] = a0;
        p1[(size_t)y * dec_dim + o] = a1;
      }
    sm.conv_p0 = upload(p0);
    sm.conv_p1 = upload(p1);
  }
  {
    const double a = 1.0 / 3.0;
    sm.ts_max = V > 1 ? (1.0 / (a - 1.0)) * (1.0 - pow((double)V, 1.0 - a)) : 1.0;
    sm.max_ent = V > 1 ? log((double)V) : 1.0;
  }

  // tokens
  const std::string tp = str(mc.tokens);
  id2token.assign(V, "");
  if (!tp.empty()) {
    std::ifstream tf(tp);
    if (!tf) throw std::runtime_error("cannot open tokens file: " + tp);
    std::string ln;
    while (std::getline(tf, ln)) {
      const size_t sp2 = ln.find_last_of(" \t");
      if (sp2 == std::string::npos) continue;
      const int id = atoi(ln.c_str() + sp2 + 1);
      std::string sym = ln.substr(0, sp2);
      while (!sym.empty() && (sym.back() == ' ' || sym.back() == '\t')) sym.pop_back();
      if (id >= 0 && id < V) id2token[id] = sym;
    }
  }
  decoding_method = str(c->decoding_method).empty() ? "greedy_search" : str(c->decoding_method);
  if (decoding_method != "greedy_search" && decoding_method != "modified_beam_search")
    throw std::runtime_error("unsupported decoding_method: " + decoding_method);
  max_active_paths = c->max_active_paths > 0 ? c->max_active_paths : 4;
  if (max_active_paths > 16) throw std::runtime_error("max_active_paths > 16 is not built");
  hotwords_score = c->hotwords_score > 0 ? c->hotwords_score : 1.5f;
  blank_penalty = c->blank_penalty;
  fbank_tables_create(&fb);
  search = search_state_create();
  if (precision < 0 || precision > 3)
    throw std::runtime_error("precision must be 0 (FP32 mode on tcgen05), 1 (TF32 tcgen05), 2 (FP32 CUDA cores) or 3 (BF16 tcgen05)");
  // FP32 mode: the fp16 operand split (gemm_tc_f16.cu) unless B200ASR_GEMM_3XTF32 asks for the 3xTF32 kernel
  use_f16x3 = precision == 0 && getenv("B200ASR_GEMM_3XTF32") == nullptr;
  if (precision != 2 && !gemm_tc_available()) throw std::runtime_error("tensor-core GEMM path unavailable (cuTensorMapEncodeTiled not found)");
  search_set_gemm(search, gemm_fn(), precision != 2);
  if (use_f16x3 || precision == 3) {
    const W16 &w = w16_for(sm.join_w, V, join_dim);
    sm.join_w16hi = w.hi; sm.join_w16lo = w.lo; sm.join_w16_ld = w.ld;
  }
  if (precision == 0) {
    float *lo;
    CUDA_CHECK(cudaMalloc(&lo, (size_t)V * join_dim * sizeof(float)));
    owned.push_back(lo);
    launch_split_lo(sm.join_w, lo, (long long)V * join_dim, st);
    sm.join_w_lo = lo;
  }

  // hotwords file with token ids (modeling_unit token_id); text units are tokenised by the host binding
  const std::string hw = str(c->hotwords_file), mu_ = str(mc.modeling_unit);
  if (!hw.empty() && mu_ == "token_id") {
    std::ifstream hf(hw);
    if (!hf) throw std::runtime_error("cannot open hotwords file: " + hw);
    std::vector<int32_t> toks, offs{0};
    std::vector<float> scs;
    std::string ln;
    while (std::getline(hf, ln)) {
      if (ln.empty() || ln[0] == '#') continue;
      float sc = hotwords_score;
      const size_t colon = ln.rfind(':');
      if (colon != std::string::npos) { sc = (float)atof(ln.c_str() + colon + 1); ln = ln.substr(0, colon); }
      std::stringstream ls(ln);
      int id, cnt = 0;
      while (ls >> id) { toks.push_back(id); ++cnt; }
      if (cnt == 0) continue;
      offs.push_back((int32_t)toks.size());
      scs.push_back(sc);
    }
    set_graph(toks.data(), offs.data(), scs.data(), (int)scs.size());
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
}

std::shared_ptr<GraphPair> Engine::make_graph(const int32_t *tokens, const int32_t *offsets, const float *scores, int n) {
  if (n <= 0) return nullptr;
  for (int p = 0; p < n; ++p)
    for (int j = offsets[p]; j < offsets[p + 1]; ++j)
      if (tokens[j] < 0 || tokens[j] >= V) throw std::runtime_error("hotword token id out of range");
  auto g = std::make_shared<GraphPair>();
  g->host.build(tokens, offsets, scores, n);
  g->upload(device);
  return g;
}

// The stateless decoder sees only the last two tokens, so decoder_proj(relu(conv(emb[y0], emb[y1]))) takes V^2 values. With the
// whole table in HBM (V = 2000, jd = 512: 8.2 GB of the 180) a frame step of the search needs no decoder kernel at all: the
// selection kernel reads the row of each new hypothesis and writes the joiner input directly. Built once per recognizer with
// the Linear-layer GEMM of the recognizer's precision mode, in chunks of 128 K contexts (pre-activations from the per-token
// convolution tables, then one GEMM). Skipped (the search then computes decoder rows on demand) in the CUDA-core cross-check
// mode, with B200ASR_DEC_TABLE=0, or when the table would take more than a quarter of the free device memory.
// Call once the stream an encoder ran on has been synchronised. True: a softmax row left the safe range of the single-pass
// form; the caller repeats its pass (exact_softmax is set, so the repeat and everything after it use the two-pass kernel).
bool Engine::softmax_overflowed() {
  if (exact_softmax || !sm_flag_host) return false;
  if (getenv("B200ASR_DBG_FORCE_SOFTMAX_RETRY")) *sm_flag_host = 1;   // tests: take the retry path without a pathological model
  if (*sm_flag_host == 0) return false;
  *sm_flag_host = 0;
  CUDA_CHECK(cudaMemset(sm_flag, 0, sizeof(int)));
  exact_softmax = true;
  return true;
}

void Engine::ensure_dec_table() {
  if (dec_table_tried) return;
  dec_table_tried = true;
  if (precision == 2) return;
  if (const char *e = getenv("B200ASR_DEC_TABLE")) if (atoi(e) == 0) return;
  const size_t n_ctx = (size_t)V * V;
  const size_t bytes = n_ctx * join_dim * sizeof(float);
  size_t free_b = 0, total_b = 0;
  CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  const int chunk = 128 * 1024;
  if (bytes + (size_t)chunk * dec_dim * sizeof(float) > free_b / 4) return;
  float *table = nullptr, *pre = nullptr;
  if (cudaMalloc(&table, bytes) != cudaSuccess) { cudaGetLastError(); return; }
  owned.push_back(table);
  CUDA_CHECK(cudaMalloc(&pre, (size_t)chunk * dec_dim * sizeof(float)));
  const double f0 = gemm_flops, b0 = gemm_bytes;
  const long long n0 = gemm_launches;
  const bool prof = profiling;
  profiling = false;
  reset_tile_counters();   // the GEMMs below claim tiles through counters that must be zero at launch
  for (size_t c0 = 0; c0 < n_ctx; c0 += chunk) {
    const int rows = (int)std::min<size_t>(chunk, n_ctx - c0);
    launch_context_preactivations(sm, (long long)c0, rows, pre, st);
    gemm(pre, dec_dim, sm.dec_proj_w, sm.dec_proj_b, nullptr, 0, table + c0 * join_dim, join_dim, rows, join_dim, dec_dim, ACT_NONE);
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  CUDA_CHECK(cudaFree(pre));
  profiling = prof;
  gemm_flops = f0; gemm_bytes = b0; gemm_launches = n0;
  sm.dec_table = table;
}

void Engine::set_graph(const int32_t *tokens, const int32_t *offsets, const float *scores, int n) {
  graph = make_graph(tokens, offsets, scores, n);
}

void Engine::gemm(const float *A, int lda, const float *Wt, const float *bias, const float *R, int ldr, float *C, int ldc, int M,
                  int N, int K, int act) {
  static const bool hprof = getenv("B200ASR_HOST_PROF") != nullptr;
  struct Tm { double *acc; std::chrono::steady_clock::time_point t0; ~Tm() { if (acc) *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } };
  Tm _tm{hprof ? &hp_gemm : nullptr, std::chrono::steady_clock::now()};
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = Wt; g.bias = bias; g.R = R; g.ldr = ldr; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.act = act;
  // programmatic dependent launch: the kernel's prologue (barrier init, TMEM allocation, tensor-map prefetch) runs on SMs
  // the previous kernel has already left; it waits for that kernel's completion before its first global access
  static const bool enc_pdl = getenv("B200ASR_NO_PDL") == nullptr;
  g.pdl = (enc_pdl && !profiling) ? 1 : 0;
  if (M <= 0) return;
  g.tile_counter = next_tile_counter();
  if (precision == 0) {
    auto it = w_lo.find(Wt);
    if (it == w_lo.end()) {   // first use of this weight: split once, keep
      float *lo;
      CUDA_CHECK(cudaMalloc(&lo, (size_t)N * K * sizeof(float)));
      owned.push_back(lo);
      launch_split_lo(Wt, lo, (long long)N * K, st);
      it = w_lo.emplace(Wt, lo).first;
    }
    g.Wlo = it->second;
  }
  if (use_f16x3 || precision == 3) {
    const W16 &w = w16_for(Wt, N, K);
    g.W16hi = w.hi; g.W16lo = w.lo; g.w16_ld = w.ld;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (profiling) {
    if (gemm_ev_used >= gemm_events.size()) {
      cudaEvent_t a, b;
      CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
      gemm_events.emplace_back(a, b);
    }
    e0 = gemm_events[gemm_ev_used].first; e1 = gemm_events[gemm_ev_used].second;
    ++gemm_ev_used;
    CUDA_CHECK(cudaEventRecord(e0, st));
  }
  gemm_fn()(g, st);
  if (profiling) CUDA_CHECK(cudaEventRecord(e1, st));
  gemm_flops += 2.0 * (double)M * (double)N * (double)K;
  gemm_bytes += 4.0 * ((double)M * K + (double)N * K + (double)M * N * (R ? 2.0 : 1.0));   // operands once, result (+ residual)
  ++gemm_launches;
}

void Engine::collect_gemm_times() {
  gemm_ms = 0;
  for (size_t i = 0; i < gemm_ev_used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, gemm_events[i].first, gemm_events[i].second) == cudaSuccess) gemm_ms += ms;
  }
  gemm_ev_used = 0;
}

// ------------------------------------------------------------------ fbank
void Engine::run_fbank(const float *d_pcm, const long long *d_soff, const long long *d_slen, const std::vector<long long> &h_len,
                       int n, float **d_feats, std::vector<int> *T) {
  std::vector<long long> foff(n + 1, 0);
  T->assign(n, 0);
  int maxT = 0;
  for (int u = 0; u < n; ++u) {
    const long long ns = h_len[u];
    (*T)[u] = (int)((ns + 80) / 160);
    foff[u + 1] = foff[u] + (*T)[u];
    maxT = std::max(maxT, (*T)[u]);
  }
  long long *d_foff = b_foff.get<long long>(n + 1);
  const long long total_frames = foff[n];
  CUDA_CHECK(cudaMemcpyAsync(d_foff, keep(std::move(foff)), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  float *feats = b_feats.get<float>((size_t)std::max<long long>(total_frames, 1) * 80);
  launch_fbank(fb, d_pcm, d_soff, d_slen, d_foff, n, maxT, feats, st);
  *d_feats = feats;
}

// ------------------------------------------------------------------ encoder
void Engine::build_attn_plan(AttnPlan *pl, const StackW &s, int q, int n, const std::vector<long long> &aoff_host) {
  static const bool hprof = getenv("B200ASR_HOST_PROF") != nullptr;
  struct Tm { double *acc; std::chrono::steady_clock::time_point t0; ~Tm() { if (acc) *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } };
  Tm _tm{hprof ? &hp_plan : nullptr, std::chrono::steady_clock::now()};
  *pl = AttnPlan{};
  if (precision == 2 || n <= 0) return;   // CUDA-core mode keeps the CUDA-core attention kernels
  const int H = s.H, C12 = H * vd, hid = (3 * s.D) / 4;
  const std::vector<int> &len = h_len[q];
  std::vector<long long> vt12(n + 1, 0), vth(n + 1, 0);
  std::vector<int> t12(n + 1, 0), th(n + 1, 0);
  for (int u = 0; u < n; ++u) {
    const int Tk = len[u], Tk4 = (Tk + 3) & ~3, mt = (Tk + 127) / 128;
    vt12[u + 1] = vt12[u] + (long long)C12 * Tk4;
    vth[u + 1] = vth[u] + (long long)hid * Tk4;
    t12[u + 1] = t12[u] + H * mt;
    th[u + 1] = th[u] + mt * ((hid + 63) / 64);
  }
  const bool split3 = precision == 0;
  pl->VT12 = b_vt12.get<float>((size_t)std::max<long long>(vt12[n], 4));
  pl->VTh = b_vth.get<float>((size_t)std::max<long long>(vth[n], 4));
  if (split3) {
    pl->VT12lo = b_vt12lo.get<float>((size_t)std::max<long long>(vt12[n], 4));
    pl->VThlo = b_vthlo.get<float>((size_t)std::max<long long>(vth[n], 4));
  }
  // tensor maps (128 bytes each): A, V12, V12lo, Vh, Vhlo
  unsigned char *hp = pinned_scratch((size_t)5 * n * 128);
  if (!hp) {
    unsigned char *hm = const_cast<unsigned char *>(keep(std::vector<unsigned char>((size_t)5 * n * 128 + 64)));
    hp = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(hm) + 63) & ~uintptr_t(63));
  }
  // 5 n driver calls (cuTensorMapEncodeTiled, ~3 us each): 26 of the 27 ms of host time one encoder pass of the bench batch took
  // were these, which in chained passes delayed the launch sequence of the previous batch's search by as much. The encodes
  // are independent host work, so they are dealt in blocks of 32 utterances to a few threads.
  {
    struct Job { int which, u0, u1; };
    std::vector<Job> jobs;
    for (int w = 0; w < 5; ++w)
      for (int u0 = 0; u0 < n; u0 += 32) jobs.push_back(Job{w, u0, std::min(n, u0 + 32)});
    const float *bases[5] = {b_A.ptr<float>(), pl->VT12, split3 ? pl->VT12lo : pl->VT12, pl->VTh, split3 ? pl->VThlo : pl->VTh};
    const long long *offs[5] = {aoff_host.data(), vt12.data(), vt12.data(), vth.data(), vth.data()};
    const int rows_mult[5] = {H, 0, 0, 0, 0}, rows_fixed[5] = {0, C12, C12, hid, hid}, box_rows[5] = {128, 16, 16, 64, 64};
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex err_mu;
    auto work = [&] {
      try {
        CUDA_CHECK(cudaSetDevice(device));     // the driver call wants this device's context current in the calling thread
        for (size_t j = next.fetch_add(1); j < jobs.size(); j = next.fetch_add(1)) {
          const Job &jb = jobs[j];
          attn_tc_encode_maps(hp + ((size_t)jb.which * n + jb.u0) * 128, jb.u1 - jb.u0, bases[jb.which], offs[jb.which] + jb.u0,
                              len.data() + jb.u0, rows_mult[jb.which], rows_fixed[jb.which], box_rows[jb.which]);
        }
      } catch (...) {
        std::lock_guard<std::mutex> lk(err_mu);
        if (!err) err = std::current_exception();
      }
    };
    const int helpers = std::min<int>(host_helper_threads(), (int)jobs.size() / 4);
    std::vector<std::thread> ths;
    for (int i = 0; i < helpers; ++i) ths.emplace_back(work);
    work();
    for (auto &t : ths) t.join();
    if (err) std::rethrow_exception(err);
  }
  unsigned char *dm = b_maps.get<unsigned char>((size_t)5 * n * 128);
  CUDA_CHECK(cudaMemcpyAsync(dm, hp, (size_t)5 * n * 128, cudaMemcpyHostToDevice, st));
  const int nt12 = t12[n], nth = th[n];
  int *dt = b_tileoff.get<int>((size_t)2 * (n + 1));
  CUDA_CHECK(cudaMemcpyAsync(dt, keep(std::move(t12)), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(dt + (n + 1), keep(std::move(th)), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  long long *dv = b_vtoff.get<long long>((size_t)2 * (n + 1));
  CUDA_CHECK(cudaMemcpyAsync(dv, keep(std::move(vt12)), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(dv + (n + 1), keep(std::move(vth)), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  pl->mapsA = dm; pl->mapsV12 = dm + 1 * (size_t)n * 128; pl->mapsV12lo = dm + 2 * (size_t)n * 128;
  pl->mapsVh = dm + 3 * (size_t)n * 128; pl->mapsVhlo = dm + 4 * (size_t)n * 128;
  pl->tile_off12 = dt; pl->tile_offh = dt + (n + 1);
  pl->vt_off12 = dv; pl->vt_offh = dv + (n + 1);
  pl->n_tiles12 = nt12; pl->n_tilesh = nth;
  pl->use = true;
}

void Engine::attn_apply(const AttnPlan &pl, const StackW &s, const RaggedDesc &r, const long long *aoff, const float *X, int ldx,
                        const float *S, int lds, const float *Y, int ldy, int C, bool single_head, float *out, int ldo, const float *Ls) {
  if (!pl.use) {
    launch_attn_apply(b_A.ptr<float>(), aoff, r, X, ldx, S, lds, Y, ldy, C, vd, single_head ? 1 : 0, out, ldo, st);
    return;
  }
  const bool split3 = precision == 0;
  AttnTcLaunch a{};
  a.mapsA = pl.mapsA; a.len = r.len; a.off = r.off; a.n_utt = r.n; a.single_head = single_head ? 1 : 0; a.C = C; a.dv = vd;
  a.Y = Y; a.ldy = ldy; a.out = out; a.ldo = ldo; a.split3 = split3 ? 1 : 0;
  a.tile_counter = next_tile_counter();
  a.Ls = Ls; a.H = s.H;
  if (single_head) {
    launch_transpose_v(X, ldx, S, lds, C, r, pl.dw_tile_off, pl.dw_tiles, pl.vt_offh, pl.VTh, split3 ? pl.VThlo : nullptr, st);
    a.mapsV = pl.mapsVh; a.mapsVlo = pl.mapsVhlo; a.tile_off = pl.tile_offh; a.n_tiles = pl.n_tilesh;
  } else {
    launch_transpose_v(X, ldx, S, lds, C, r, pl.dw_tile_off, pl.dw_tiles, pl.vt_off12, pl.VT12, split3 ? pl.VT12lo : nullptr, st);
    a.mapsV = pl.mapsV12; a.mapsVlo = pl.mapsV12lo; a.tile_off = pl.tile_off12; a.n_tiles = pl.n_tiles12;
  }
  launch_attn_apply_tc(a, st);
}

// out = R + Linear(SwooshL(Linear(x))). The hidden activation [M, f] is the largest tensor of a layer (up to 5 D wide); walking
// the rows in chunks whose hidden block stays in L2 keeps it out of HBM: the second GEMM of a chunk reads what the first has
// just written, and the next chunk overwrites the same lines before they are evicted (B200ASR_FFN_CHUNK_MB, 0 = one pass).
void Engine::feed_forward(const LayerW &w, int j, const float *x, const float *R, float *out, int M, int D, float *hidden) {
  const int f = w.ff_dim[j];
  static const long long chunk_mb = getenv("B200ASR_FFN_CHUNK_MB") ? atoll(getenv("B200ASR_FFN_CHUNK_MB")) : 0;
  long long rows = M;
  if (chunk_mb > 0) {
    rows = std::max<long long>(128 * 148, ((chunk_mb << 20) / ((long long)f * 4)) / 128 * 128);
    if (rows * 5 / 4 >= M) rows = M;      // no sliver at the end
  }
  for (long long m0 = 0; m0 < M; m0 += rows) {
    const int mc = (int)std::min<long long>(rows, M - m0);
    gemm(x + m0 * D, D, w.ff_in_w[j], w.ff_in_b[j], nullptr, 0, hidden, f, mc, f, D, ACT_SWOOSH_L);
    gemm(hidden, f, w.ff_out_w[j], w.ff_out_b[j], R + m0 * D, D, out + m0 * D, D, mc, D, f, ACT_NONE);
  }
}

void Engine::run_layer(const StackW &s, const LayerW &w, float *src, const RaggedDesc &r, const long long *aoff, int M, int Lmax,
                       const AttnPlan &pl) {
  const int D = s.D, H = s.H, h = (3 * D) / 4;
  const int pw = H * (2 * qd + pd);
  int maxw = std::max({pw, 3 * h, 2 * D, w.ff_dim[2], H * vd});
  float *proj = b_proj.get<float>((size_t)M * maxw);
  float *hid = b_hid.get<float>((size_t)M * std::max({h, D, H * vd}));
  float *w1 = b_w1.get<float>((size_t)M * D);
  float *A = b_A.ptr<float>();
  float *pp = b_pp.get<float>((size_t)(2 * Lmax - 1) * H * pd);
  // attention weights (computed once per layer from the layer input)
  gemm(src, D, w.attn_in_w, w.attn_in_b, nullptr, 0, proj, pw, M, pw, D, ACT_NONE);
  gemm(b_pe.ptr<float>(), pos_dim, w.pos_w, nullptr, nullptr, 0, pp, H * pd, 2 * Lmax - 1, H * pd, pos_dim, ACT_NONE);
  float *Ls = nullptr;
  if (pl.use && attn_weights_tc_supported(qd, pd) && !getenv("B200ASR_ATTN_SIMT")) {
    if (!exact_softmax) Ls = b_ls.get<float>((size_t)M * H);
    launch_attn_weights_tc(proj, pw, M, pp, r, aoff, pl.tile_off12, pl.n_tiles12, H, A, precision == 0, st, next_tile_counter(), Ls, sm_flag);
  } else {
    launch_attn_weights(proj, pw, pp, r, aoff, H, qd, pd, A, st);
  }
  // feed_forward1
  feed_forward(w, 0, src, src, w1, M, D, proj);
  // nonlin attention (head 0)
  gemm(w1, D, w.nl_in_w, w.nl_in_b, nullptr, 0, proj, 3 * h, M, 3 * h, D, ACT_NONE);
  attn_apply(pl, s, r, aoff, proj + h, 3 * h, proj, 3 * h, proj + 2 * h, 3 * h, h, true, hid, h, Ls);
  gemm(hid, h, w.nl_out_w, w.nl_out_b, w1, D, w1, D, M, D, h, ACT_NONE);
  for (int j = 0; j < 2; ++j) {
    // self attention j
    gemm(w1, D, w.sa_in_w[j], w.sa_in_b[j], nullptr, 0, proj, H * vd, M, H * vd, D, ACT_NONE);
    attn_apply(pl, s, r, aoff, proj, H * vd, nullptr, 0, nullptr, 0, H * vd, false, hid, H * vd, Ls);
    gemm(hid, H * vd, w.sa_out_w[j], w.sa_out_b[j], w1, D, w1, D, M, D, H * vd, ACT_NONE);
    // conv module j
    gemm(w1, D, w.cv_in_w[j], w.cv_in_b[j], nullptr, 0, proj, 2 * D, M, 2 * D, D, ACT_NONE);
    launch_glu_dwconv(proj, r, pl.dw_tile_off, pl.dw_tiles, D, s.k, w.cv_dw_w[j], w.cv_dw_b[j], hid, st);
    gemm(hid, D, w.cv_out_w[j], w.cv_out_b[j], w1, D, w1, D, M, D, D, ACT_NONE);
    // feed forward 2 / 3
    feed_forward(w, j + 1, w1, w1, w1, M, D, proj);
    if (j == 0) launch_bypass(w1, src, M, D, w.bypass_mid, w1, st);
  }
  launch_biasnorm_bypass(w1, src, M, D, w.norm_bias, w.norm_log_scale, w.bypass, src, st);
}

void Engine::run_encoder(const float *d_feats, const std::vector<int> &T, DevBuf &enc_buf, float **d_enc, std::vector<int> *Tp) {
  const int n = (int)T.size();
  last_n = n;
  h_T = T;
  h_T1.assign(n, 0);
  std::vector<long long> foff(n + 1, 0), c0off(n + 1, 0), c1off(n + 1, 0);
  int maxT = 0, maxt2 = 0;
  for (int u = 0; u < n; ++u) {
    const int t = T[u];
    h_T1[u] = t >= 9 ? (t - 7) / 2 : 0;
    const int t2 = t >= 9 ? (t - 5) / 2 + 1 : 0;
    foff[u + 1] = foff[u] + t;
    c0off[u + 1] = c0off[u] + (t >= 9 ? t - 2 : 0);
    c1off[u + 1] = c1off[u] + t2;
    if (t >= 9) { maxT = std::max(maxT, t); maxt2 = std::max(maxt2, t2); }
  }
  const int rates[4] = {1, 2, 4, 8};
  int Mr[4], Lr[4];
  for (int q = 0; q < 4; ++q) {
    h_len[q].assign(n, 0); h_off[q].assign(n + 1, 0);
    Lr[q] = 0;
    for (int u = 0; u < n; ++u) {
      h_len[q][u] = (h_T1[u] + rates[q] - 1) / rates[q];
      h_off[q][u + 1] = h_off[q][u] + h_len[q][u];
      Lr[q] = std::max(Lr[q], h_len[q][u]);
    }
    Mr[q] = h_off[q][n];
  }
  Tp->assign(h_len[1].begin(), h_len[1].end());   // T' = (T1+1)//2
  h_Tp = *Tp;
  const int M1 = Mr[0];
  // upload descriptors
  int *d_T = b_T.get<int>(n);
  std::vector<int> Tclamped(n);
  for (int u = 0; u < n; ++u) Tclamped[u] = T[u] >= 9 ? T[u] : 0;
  CUDA_CHECK(cudaMemcpyAsync(d_T, Tclamped.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  long long *d_foff = b_foff.get<long long>(n + 1);
  CUDA_CHECK(cudaMemcpyAsync(d_foff, foff.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  long long *d_c0off = b_c0off.get<long long>(n + 1), *d_c1off = b_c1off.get<long long>(n + 1);
  CUDA_CHECK(cudaMemcpyAsync(d_c0off, c0off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_CHECK(cudaMemcpyAsync(d_c1off, c1off.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  RaggedDesc rd[4];
  for (int q = 0; q < 4; ++q) {
    int *dl = b_len[q].get<int>(n), *dof = b_off[q].get<int>(n + 1);
    CUDA_CHECK(cudaMemcpyAsync(dl, h_len[q].data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(dof, h_off[q].data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    rd[q] = RaggedDesc{dl, dof, n, Mr[q], Lr[q]};
  }
  // row maps between the full rate and each stack rate; 128-frame tile lists per rate (conv module)
  int *d_up[4] = {nullptr, nullptr, nullptr, nullptr};
  int2 *d_down[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int q = 1; q < 4; ++q) {
    d_up[q] = b_up[q].get<int>((size_t)std::max(Mr[0], 1));
    d_down[q] = b_down[q].get<int2>((size_t)std::max(Mr[q], 1));
    launch_build_row_maps(rd[0], rd[q], rates[q], d_up[q], d_down[q], st);
  }
  std::vector<int> dwt((size_t)4 * (n + 1), 0);
  for (int q = 0; q < 4; ++q)
    for (int u = 0; u < n; ++u) dwt[(size_t)q * (n + 1) + u + 1] = dwt[(size_t)q * (n + 1) + u] + (h_len[q][u] + 127) / 128;
  int *d_dwt = b_dwt.get<int>(dwt.size());
  CUDA_CHECK(cudaMemcpyAsync(d_dwt, dwt.data(), dwt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  // attention-weight offsets per stack and the max A size
  const size_t ns = stacks.size();
  std::vector<std::vector<long long>> aoffs(ns, std::vector<long long>(n + 1, 0));
  long long maxA = 1;
  auto rate_idx = [](int ds) { return ds == 1 ? 0 : ds == 2 ? 1 : ds == 4 ? 2 : 3; };
  for (size_t i = 0; i < ns; ++i) {
    const int q = rate_idx(stacks[i].ds);
    for (int u = 0; u < n; ++u)
      aoffs[i][u + 1] = aoffs[i][u] + (long long)stacks[i].H * h_len[q][u] * ((h_len[q][u] + 3) & ~3);
    maxA = std::max(maxA, aoffs[i][n]);
  }
  long long *d_aoff = b_aoff.get<long long>(ns * (n + 1));
  for (size_t i = 0; i < ns; ++i)
    CUDA_CHECK(cudaMemcpyAsync(d_aoff + i * (n + 1), aoffs[i].data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  b_A.get<float>((size_t)maxA);
  float *enc = enc_buf.get<float>((size_t)std::max(Mr[1], 1) * join_dim);
  *d_enc = enc;
  if (M1 <= 0) {
    keep(std::move(foff)); keep(std::move(c0off)); keep(std::move(c1off)); keep(std::move(Tclamped)); keep(std::move(dwt));
    for (auto &v : aoffs) keep(std::move(v));
    return;
  }

  // ---- Conv2dSubsampling
  const int D0 = enc_dim[0];
  float *c0 = b_c0.get<float>((size_t)c0off[n] * 640);
  float *c1 = b_c1.get<float>((size_t)c1off[n] * 39 * 32);
  float *c2 = b_c2.get<float>((size_t)M1 * 2432);
  float *dw = b_dw.get<float>((size_t)M1 * 2432);
  float *pw1 = b_pw1.get<float>((size_t)M1 * 19 * 384);
  float *cn = b_cn.get<float>((size_t)M1 * 2432);
  float *x0 = b_x0.get<float>((size_t)M1 * D0);
  launch_embed_conv0(d_feats, d_foff, d_c0off, n, c0off[n], w_conv0, W("encoder.embed.conv0.bias"), c0, st);
  launch_embed_conv1(c0, d_c0off, d_c1off, n, c1off[n], w_conv1, W("encoder.embed.conv1.bias"), c1, st);
  launch_embed_im2col2(c1, d_c1off, rd[0].off, n, M1, pw1, st);   // pw1 buffer doubles as the im2col scratch (288 <= 384)
  gemm(pw1, 288, w_conv2, W("encoder.embed.conv2.bias"), nullptr, 0, c2, 128, M1 * 19, 128, 288, ACT_SWOOSH_R);
  launch_embed_dw7(c2, rd[0], d_dwt, dwt[n], w_dw7, W("encoder.embed.convnext.dw.bias"), dw, st);
  {
    // ConvNeXt pointwise pair over T1 * 19 rows: the 384-wide hidden tensor (4.2 GB on the bench batch) is three times the size
    // of its input and output, and both GEMMs are HBM-bound (K = 128 / N = 128). Walking the rows in chunks whose hidden block
    // stays in L2 keeps it out of HBM (B200ASR_EMBED_CHUNK_MB, 0 = one pass).
    static const long long chunk_mb = getenv("B200ASR_EMBED_CHUNK_MB") ? atoll(getenv("B200ASR_EMBED_CHUNK_MB")) : 0;
    const long long rows_all = (long long)M1 * 19;
    long long rows_c = rows_all;
    if (chunk_mb > 0) {
      rows_c = std::max<long long>(128 * 148, ((chunk_mb << 20) / (384 * 4)) / (128 * 148) * (128 * 148));
      if (rows_c * 5 / 4 >= rows_all) rows_c = rows_all;
    }
    const float *w1 = W("encoder.embed.convnext.pw1.weight"), *b1 = W("encoder.embed.convnext.pw1.bias");
    const float *w2 = W("encoder.embed.convnext.pw2.weight"), *b2 = W("encoder.embed.convnext.pw2.bias");
    for (long long m0 = 0; m0 < rows_all; m0 += rows_c) {
      const int mc = (int)std::min<long long>(rows_c, rows_all - m0);
      gemm(dw + m0 * 128, 128, w1, b1, nullptr, 0, pw1, 384, mc, 384, 128, ACT_SWOOSH_L);
      gemm(pw1, 384, w2, b2, c2 + m0 * 128, 128, cn + m0 * 128, 128, mc, 128, 384, ACT_NONE);
    }
  }
  gemm(cn, 2432, w_out, W("encoder.embed.out.bias"), nullptr, 0, dw, D0, M1, D0, 2432, ACT_NONE);   // dw reused as scratch
  launch_biasnorm(dw, M1, D0, W("encoder.embed.out_norm.bias"), W("encoder.embed.out_norm.log_scale"), x0, st);

  // ---- stacks
  const float *x = x0;
  int xC = D0;
  for (size_t i = 0; i < ns; ++i) {
    const StackW &s = stacks[i];
    const int q = rate_idx(s.ds);
    const int D = s.D;
    float *outb = b_stack[i].get<float>((size_t)M1 * D);
    const long long *aoff = d_aoff + i * (n + 1);
    float *pe = b_pe.get<float>((size_t)(2 * Lr[q] - 1) * pos_dim);
    launch_pos_emb(pe, Lr[q], pos_dim, st);
    AttnPlan plan;
    build_attn_plan(&plan, s, q, n, aoffs[i]);
    plan.dw_tile_off = d_dwt + (size_t)q * (n + 1);
    plan.dw_tiles = dwt[(size_t)q * (n + 1) + n];
    if (s.ds == 1) {
      launch_convert_channels(x, xC, outb, D, M1, st);
      for (int l = 0; l < s.L; ++l) run_layer(s, s.layers[l], outb, rd[0], aoff, M1, Lr[0], plan);
    } else {
      float *xc = b_xc.get<float>((size_t)M1 * D);
      launch_convert_channels(x, xC, xc, D, M1, st);
      float *sin_ = b_sin.get<float>((size_t)Mr[q] * D);
      launch_downsample(xc, d_down[q], Mr[q], D, s.ds, s.ds_bias, sin_, st);
      for (int l = 0; l < s.L; ++l) run_layer(s, s.layers[l], sin_, rd[q], aoff, Mr[q], Lr[q], plan);
      launch_upsample_combine(sin_, d_up[q], xc, M1, D, s.combiner, outb, st);
    }
    x = outb;
    xC = D;
  }
  // ---- full-dim output: each channel range from the latest stack that has it, then downsample(2), encoder_proj
  std::vector<ConcatPiece> pieces;
  int curd = enc_dim[ns - 1];
  pieces.push_back(ConcatPiece{b_stack[ns - 1].ptr<float>(), enc_dim[ns - 1], 0, curd});
  for (int i = (int)ns - 2; i >= 0; --i) {
    if (enc_dim[i] > curd) {
      pieces.push_back(ConcatPiece{b_stack[i].ptr<float>(), enc_dim[i], curd, enc_dim[i]});
      curd = enc_dim[i];
    }
  }
  if (pieces.size() > 4) throw std::runtime_error("more than 4 output pieces not built");
  float *cat = b_cat.get<float>((size_t)Mr[1] * out_dim);
  launch_concat_downsample2(pieces.data(), (int)pieces.size(), d_down[1], Mr[1], out_dim, W("encoder.downsample_output.bias"), cat, st);
  gemm(cat, out_dim, W("encoder.encoder_proj.weight"), W("encoder.encoder_proj.bias"), nullptr, 0, enc, join_dim, Mr[1], join_dim,
       out_dim, ACT_NONE);
  if (!exact_softmax) CUDA_CHECK(cudaMemcpyAsync(sm_flag_host, sm_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  // the descriptor vectors uploaded asynchronously above stay alive until the next pass (no synchronisation here)
  keep(std::move(foff)); keep(std::move(c0off)); keep(std::move(c1off)); keep(std::move(Tclamped)); keep(std::move(dwt));
  for (auto &v : aoffs) keep(std::move(v));
}

// ------------------------------------------------------------------ full pipeline on device-resident PCM
void Engine::ensure_partition() {
  if (part.tried) return;
  part.tried = true;
  const GreenApi &ga = green_api();
  if (!ga.ok) return;
  static const int want = getenv("B200ASR_SM_RESERVE") ? atoi(getenv("B200ASR_SM_RESERVE")) : 16;
  // Off unless asked for (B200ASR_GREEN_CTX=1). Measured on a B200 (C2, 4 groups): the search's step kernels want 60-100 SMs
  // for a few microseconds at a time; confined to a 16-SM partition a frame step takes 185 us instead of 37 (32 SMs: 100 us)
  // and the pass gets slower, not faster. Stream priority plus SMs left free by the encoder's persistent grids does better.
  static const bool green = getenv("B200ASR_GREEN_CTX") != nullptr;
  if (!green || want <= 0) return;
  CUdevice dev;
  CUdevResource sm{}, grp[1]{}, rem{};
  unsigned int n = 1;
  CUdevResourceDesc d_search = nullptr, d_enc = nullptr;
  auto ok = [](CUresult r) { return r == CUDA_SUCCESS; };
  if (!ok(ga.DeviceGet(&dev, device)) || !ok(ga.DeviceGetDevResource(dev, &sm, CU_DEV_RESOURCE_TYPE_SM))) return;
  if (!ok(ga.DevSmResourceSplitByCount(grp, &n, &sm, &rem, 0, (unsigned)want)) || n < 1 || rem.sm.smCount < 64) return;
  if (!ok(ga.DevResourceGenerateDesc(&d_search, &grp[0], 1)) || !ok(ga.DevResourceGenerateDesc(&d_enc, &rem, 1))) return;
  if (!ok(ga.GreenCtxCreate(&part.g_search, d_search, dev, CU_GREEN_CTX_DEFAULT_STREAM))) return;
  if (!ok(ga.GreenCtxCreate(&part.g_enc, d_enc, dev, CU_GREEN_CTX_DEFAULT_STREAM))) { ga.GreenCtxDestroy(part.g_search); part.g_search = nullptr; return; }
  CUstream s = nullptr;
  if (!ok(ga.GreenCtxStreamCreate(&s, part.g_enc, CU_STREAM_NON_BLOCKING, 0))) {
    ga.GreenCtxDestroy(part.g_search); ga.GreenCtxDestroy(part.g_enc);
    part.g_search = part.g_enc = nullptr;
    return;
  }
  st_part = reinterpret_cast<cudaStream_t>(s);
  CUDA_CHECK(cudaEventCreateWithFlags(&ev_part, cudaEventDisableTiming));
  part.search_sms = (int)grp[0].sm.smCount;
  part.enc_sms = (int)rem.sm.smCount;
  part.ok = true;
  if (getenv("B200ASR_DEBUG"))
    fprintf(stderr, "[b200asr] SM partition: %d SMs for the search lanes, %d for the encoder\n", part.search_sms, part.enc_sms);
}

Lane &Engine::lane(int i) {
  Lane &l = lanes[i];
  if (!l.st) {
    ensure_partition();
    int lo = 0, hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = greatest priority
    CUstream gs = nullptr;
    if (part.ok && green_api().GreenCtxStreamCreate(&gs, part.g_search, CU_STREAM_NON_BLOCKING, hi) == CUDA_SUCCESS)
      l.st = reinterpret_cast<cudaStream_t>(gs);
    else
      CUDA_CHECK(cudaStreamCreateWithPriority(&l.st, cudaStreamNonBlocking, hi));
    cudaEvent_t *evs[] = {&l.f0, &l.f1, &l.e1, &l.s0, &l.s1};
    for (cudaEvent_t *e : evs) CUDA_CHECK(cudaEventCreate(e));
    if (i == 0) l.search = search;
    else {
      l.search = search_state_create();
      search_set_gemm(l.search, gemm_fn(), precision != 2);
    }
  }
  return l;
}

// Split of a batch into groups that are encoded one after the other (longest utterances first) while the searches of the
// groups already encoded run beside the encoder on their own streams. The search is T' dependent frame steps - its time
// is set by the longest utterance of its group, not by the group's size - so: the last group holds the shortest
// utterances (its search is the only one nothing can hide), and going backwards each earlier group may hold utterances
// longer by what the encoder time of the groups after it can cover: L(g-1) <= L(g) + audio(g) / kappa, where
// kappa = (search ms per second of utterance length) / (encoder ms per audio-second), measured on the previous passes.
// B200ASR_PIPELINE=0 turns the split off; B200ASR_GROUPS="27,100,78" fixes the group sizes (longest first, rest last).
std::vector<std::vector<int>> Engine::plan_groups(const std::vector<long long> &h_len) const {
  const int n = (int)h_len.size();
  std::vector<std::vector<int>> single(1);
  single[0].resize(n);
  for (int u = 0; u < n; ++u) single[0][u] = u;
  // Off unless B200ASR_PIPELINE=1. Measured on a B200 (C2, Zipformer-68M): the search's step kernels are wide and shallow - a
  // frame step spreads 64-300 CTAs of 100-200 KB shared memory over the SMs for a few microseconds - so beside them the
  // encoder's 200 KB CTAs find about half the SMs free; splitting the encoder into g groups also costs ~3.5 ms of launch-bound
  // small kernels per group. 4 groups: 86 ms against 80 ms for the plain pass. Kept for a search kernel that fits few SMs.
  static const bool off = !(getenv("B200ASR_PIPELINE") && atoi(getenv("B200ASR_PIPELINE")) != 0);
  if (off || profiling || n < 16) return single;
  std::vector<int> asc(n);
  for (int u = 0; u < n; ++u) asc[u] = u;
  std::stable_sort(asc.begin(), asc.end(), [&](int a, int b) { return h_len[a] < h_len[b]; });
  std::vector<std::vector<int>> rev;   // shortest group first
  if (const char *e = getenv("B200ASR_GROUPS")) {
    std::vector<int> counts = parse_int_list(e);   // longest first
    int hiu = n;
    std::vector<std::vector<int>> fwd;
    for (int c : counts) {
      if ((int)fwd.size() + 1 >= kMaxLanes || c <= 0 || hiu <= 0) break;
      const int lo = std::max(0, hiu - c);
      fwd.emplace_back(asc.begin() + lo, asc.begin() + hiu);
      hiu = lo;
    }
    if (hiu > 0) fwd.emplace_back(asc.begin(), asc.begin() + hiu);
    for (auto &g : fwd) std::reverse(g.begin(), g.end());
    return fwd.empty() ? single : fwd;
  }
  double total = 0;
  for (long long v : h_len) total += (double)v / 16000.0;
  static const double min_audio = getenv("B200ASR_PIPE_MIN_AUDIO") ? atof(getenv("B200ASR_PIPE_MIN_AUDIO")) : 150.0;
  static const double kappa_env = getenv("B200ASR_PIPE_KAPPA") ? atof(getenv("B200ASR_PIPE_KAPPA")) : 0.0;
  static const int max_groups = std::min(kMaxLanes, getenv("B200ASR_PIPE_MAX_GROUPS") ? atoi(getenv("B200ASR_PIPE_MAX_GROUPS")) : 6);
  if (total < 3.0 * min_audio || max_groups < 2) return single;
  const double kap = kappa_env > 0 ? kappa_env : kappa;
  int i = 0;
  double prevL = 0, prevA = 0;
  while (i < n) {
    std::vector<int> g;
    double A = 0;
    const bool last_slot = (int)rev.size() + 1 >= max_groups;
    const double Llim = rev.empty() ? 0.0 : prevL + prevA / kap;
    while (i < n) {
      const double L = (double)h_len[asc[i]] / 16000.0;
      if (!last_slot && A >= min_audio && L > Llim) break;
      g.push_back(asc[i]);
      A += L;
      ++i;
    }
    prevL = (double)h_len[g.back()] / 16000.0;
    prevA = A;
    rev.push_back(std::move(g));
  }
  if (rev.size() >= 2) {   // a sliver of long utterances at the front would only add launches
    double A = 0;
    for (int u : rev.back()) A += (double)h_len[u] / 16000.0;
    if (A < min_audio) {
      auto tail = std::move(rev.back());
      rev.pop_back();
      rev.back().insert(rev.back().end(), tail.begin(), tail.end());
    }
  }
  if (rev.size() < 2) return single;
  std::vector<std::vector<int>> fwd(rev.rbegin(), rev.rend());
  for (auto &g : fwd) std::reverse(g.begin(), g.end());   // longest first inside a group as well
  return fwd;
}

Engine::UttResult Engine::result_of(int u) const {
  const SearchView v = search_view(lanes[lane_of[u]].search);
  const int i = idx_of[u];
  const long long o = v.off[i];
  return UttResult{v.n_tokens[i], v.tokens + o, v.frames + o, v.tok_lp + o, v.stats + 4 * o};
}

// With d_featin the pass starts from precomputed features instead of PCM: utterance u has (h_len[u] + 80) / 160 frames at
// row h_foff[u] of d_featin (h_len stays the sample count the features were computed from).
// forced_groups: the caller's own split (consecutive batches of a long decode call, or repetitions of a staged batch): group g's
// search runs on its lane beside the encoder of group g + 1 - the same mechanism as the length-sorted split, without cutting a
// batch into smaller encoder passes.
void Engine::decode_pcm_device(const float *d_pcm, const std::vector<long long> &h_soff, const std::vector<long long> &h_len, int n,
                               std::vector<int> *Tp, const float *d_featin, const std::vector<long long> *h_foff,
                               const std::vector<std::vector<int>> *forced_groups) {
  ensure_dec_table();
  gemm_flops = 0; gemm_bytes = 0; gemm_launches = 0; gemm_ev_used = 0;
  const long long l0 = g_launches;
  host_keep.clear();   // every entry point returns synchronised, so the previous pass has consumed its uploads
  pin_used = 0;
  reset_tile_counters();
  const std::vector<std::vector<int>> groups = forced_groups ? *forced_groups : plan_groups(h_len);
  const int G = (int)groups.size();
  if (G > kMaxLanes) throw std::runtime_error("too many groups in one pass");
  static const int reserve_env = getenv("B200ASR_SM_RESERVE") ? atoi(getenv("B200ASR_SM_RESERVE")) : 8;
  static const bool no_overlap = getenv("B200ASR_PIPE_NOOVERLAP") != nullptr;   // diagnostic: searches start after the last encoder
  if (G > 1) ensure_partition();
  const int reserve = part.ok ? part.search_sms : reserve_env;
  cudaStream_t const st_main = st;
  const int method = decoding_method == "greedy_search" ? 0 : 1;
  Tp->assign(n, 0);
  lane_of.assign(n, 0); idx_of.assign(n, 0);
  CUDA_CHECK(cudaEventRecord(ev[0], st));
  // Issue order: encoder of group g + 1 goes to the main stream before the (long) launch sequence of group g's search, so
  // the GPU never waits for the host between two encoders.
  auto issue_encoder = [&](int g) {
    Lane &l = lane(g);
    const std::vector<int> &mem = groups[g];
    const int ng = (int)mem.size();
    l.members = mem;
    std::vector<long long> off_len(2 * (size_t)ng + 1, 0), glen(ng);
    for (int i = 0; i < ng; ++i) {
      off_len[i] = h_soff[mem[i]];
      off_len[ng + 1 + i] = glen[i] = h_len[mem[i]];
      lane_of[mem[i]] = g; idx_of[mem[i]] = i;
    }
    // From the second group on a search runs beside the encoder: the encoder moves to its SM partition (or, without green
    // contexts, leaves `reserve` SMs out of its persistent grids). Same workspaces, so the streams are chained by events.
    if (g > 0 && part.ok && st == st_main) {
      CUDA_CHECK(cudaEventRecord(ev_part, st_main));
      CUDA_CHECK(cudaStreamWaitEvent(st_part, ev_part, 0));
      st = st_part;
    }
    set_sm_reserve(g > 0 ? reserve : 0);
    CUDA_CHECK(cudaEventRecord(l.f0, st));
    float *d_feats = nullptr, *d_enc = nullptr;
    std::vector<int> T;
    if (d_featin) {
      T.resize(ng);
      long long rows = 0;
      for (int i = 0; i < ng; ++i) { T[i] = (int)((glen[i] + 80) / 160); rows += T[i]; }
      d_feats = b_feats.get<float>((size_t)std::max<long long>(rows, 1) * 80);
      long long at = 0;
      for (int i = 0; i < ng; ++i) {
        if (T[i] > 0)
          CUDA_CHECK(cudaMemcpyAsync(d_feats + at * 80, d_featin + (*h_foff)[mem[i]] * 80, (size_t)T[i] * 80 * sizeof(float),
                                     cudaMemcpyDeviceToDevice, st));
        at += T[i];
      }
    } else {
      long long *d_sl = l.soff.get<long long>(off_len.size());
      CUDA_CHECK(cudaMemcpyAsync(d_sl, keep(std::move(off_len)), (2 * (size_t)ng + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
      run_fbank(d_pcm, d_sl, d_sl + ng + 1, glen, ng, &d_feats, &T);
    }
    CUDA_CHECK(cudaEventRecord(l.f1, st));
    run_encoder(d_feats, T, l.enc, &d_enc, &l.Tp);
    CUDA_CHECK(cudaEventRecord(l.e1, st));
    set_sm_reserve(0);
    for (int i = 0; i < ng; ++i) (*Tp)[mem[i]] = l.Tp[i];
  };
  auto issue_search = [&](int g) {
    Lane &l = lane(g);
    cudaStream_t ss = G == 1 ? st : l.st;
    if (G > 1) CUDA_CHECK(cudaStreamWaitEvent(ss, no_overlap ? lanes[G - 1].e1 : l.e1, 0));
    CUDA_CHECK(cudaEventRecord(l.s0, ss));
    l.steps = 0;
    for (int v : l.Tp) l.steps = std::max(l.steps, v);
    search_issue(l.search, sm, pass_graph, l.enc.ptr<float>(), l.Tp.data(), (int)l.members.size(), method,
                 max_active_paths, blank_penalty, ss);
    CUDA_CHECK(cudaEventRecord(l.s1, ss));
  };
  static const bool issue_prof = getenv("B200ASR_HOST_PROF") != nullptr;
  auto timed = [&](const char *what, int g, auto &&fn) {
    if (!issue_prof) { fn(); return; }
    const auto t0 = std::chrono::steady_clock::now();
    fn();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[b200asr host prof] issue %s(%d): %.2f ms of host time (gemm() %.2f, attention plans %.2f)\n", what, g, ms, hp_gemm, hp_plan);
    hp_gemm = hp_plan = 0;
  };
  timed("encoder", 0, [&] { issue_encoder(0); });
  if (no_overlap) for (int g = 1; g < G; ++g) issue_encoder(g);
  for (int g = 0; g < G; ++g) {
    if (g + 1 < G && !no_overlap) timed("encoder", g + 1, [&] { issue_encoder(g + 1); });
    timed("search", g, [&] { issue_search(g); });
  }
  if (st != st_main) {   // back to the main stream, after the partition's last encoder
    CUDA_CHECK(cudaEventRecord(ev_part, st));
    st = st_main;
    CUDA_CHECK(cudaStreamWaitEvent(st, ev_part, 0));
  }
  if (G > 1)   // the pass ends (ev[3] on the main stream) after every search
    for (int g = 0; g < G; ++g) CUDA_CHECK(cudaStreamWaitEvent(st, lanes[g].s1, 0));
  CUDA_CHECK(cudaEventRecord(ev[3], st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (softmax_overflowed()) {   // rare: repeat the pass with the exact two-pass softmax
    decode_pcm_device(d_pcm, h_soff, h_len, n, Tp, d_featin, h_foff, forced_groups);
    return;
  }
  for (int g = 0; g < G; ++g) search_print_prof(lane(g).search);
  n_groups_last = G;
  tm.fbank = tm.encoder = 0;
  search_busy_ms = 0;
  d2h_bytes_last = 0;
  double steps_ms = 0, steps_len = 0;
  for (int g = 0; g < G; ++g) {
    Lane &l = lanes[g];
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, l.f0, l.f1);
    cudaEventElapsedTime(&b, l.f1, l.e1);
    cudaEventElapsedTime(&c, l.s0, l.s1);
    tm.fbank += a; tm.encoder += b; search_busy_ms += c; lane_ms[g] = c;
    cudaEventElapsedTime(&lane_t[g][0], ev[0], l.f0); cudaEventElapsedTime(&lane_t[g][1], ev[0], l.e1);
    cudaEventElapsedTime(&lane_t[g][2], ev[0], l.s0); cudaEventElapsedTime(&lane_t[g][3], ev[0], l.s1);
    d2h_bytes_last += search_result_bytes(l.search);
    if (l.steps > 0) { steps_ms += c; steps_len += l.steps / 25.0; }
  }
  cudaEventElapsedTime(&tm.total, ev[0], ev[3]);
  tm.search = std::max(0.f, tm.total - tm.fbank - tm.encoder);   // the part of the searches no encoder hid
  // adapt the planner's ratio: search ms per second of utterance length over encoder ms per audio-second
  {
    double audio = 0;
    for (long long v : h_len) audio += (double)v / 16000.0;
    if (!forced_groups && n >= 16 && audio > 100.0 && tm.encoder > 0.f && steps_len > 0 && !profiling) {
      const double k_now = (steps_ms / steps_len) / ((double)(tm.fbank + tm.encoder) / audio);
      // hysteresis: a new plan means new workspace sizes (reallocation stalls), so only a clear change moves it
      if (k_now > 1.0 && k_now < 1000.0 && (k_now > 1.3 * kappa || k_now < 0.7 * kappa)) kappa = k_now;
    }
  }
  launches_last = g_launches - l0;
  if (profiling) collect_gemm_times();
}

static std::string json_escape(const std::string &s) {
  std::string o;
  for (unsigned char c : s) {
    if (c == '"') o += "\\\"";
    else if (c == '\\') o += "\\\\";
    else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
    else o += (char)c;
  }
  return o;
}


// sherpa-onnx style result JSON, built lazily: a batch decode should not pay for strings nobody reads
static void build_json(Stream *s) {
  if (s->json_built) return;
  const int cnt = s->res.count;
  std::string js = "{\"lang\": \"\", \"emotion\": \"\", \"event\": \"\", \"text\": \"" + json_escape(s->text) + "\", \"timestamps\": [";
  char buf[64];
  for (int j = 0; j < cnt; ++j) { snprintf(buf, sizeof buf, "%s%.3f", j ? ", " : "", s->timestamps[j]); js += buf; }
  js += "], \"tokens\": [";
  for (int j = 0; j < cnt; ++j) js += std::string(j ? ", " : "") + "\"" + json_escape(s->tok_str[j]) + "\"";
  js += "], \"ys_log_probs\": [";
  for (int j = 0; j < cnt; ++j) { snprintf(buf, sizeof buf, "%s%.6f", j ? ", " : "", s->lps[j]); js += buf; }
  js += "], \"words\": []}";
  s->json = js;
  s->res.json = s->json.c_str();
  s->json_built = true;
}

void Engine::unpack_results(Stream *const *ss, int nb, const std::vector<int> &Tp) {
  for (int i = 0; i < nb; ++i) {
      Stream *s = ss[i];
      const UttResult r = result_of(i);
      const int cnt = std::max(0, std::min(r.n_tokens, Tp[i]));
      const float dur = (float)s->n_samples() / 16000.0f;
      s->token_ids.assign(r.tokens, r.tokens + cnt);
      s->frames.assign(r.frames, r.frames + cnt);
      s->lps.assign(r.tok_lp, r.tok_lp + cnt);
      s->timestamps.resize(cnt); s->tsallis.resize(cnt); s->margin.resize(cnt); s->entropy.resize(cnt); s->top1.resize(cnt);
      s->tok_str.resize(cnt); s->tok_ptr.resize(cnt);
      s->text.clear();
      for (int j = 0; j < cnt; ++j) {
        s->timestamps[j] = Tp[i] > 0 ? (float)((double)s->frames[j] / (double)Tp[i] * (double)dur) : 0.f;
        s->tsallis[j] = r.stats[j * 4]; s->margin[j] = r.stats[j * 4 + 1]; s->entropy[j] = r.stats[j * 4 + 2]; s->top1[j] = r.stats[j * 4 + 3];
        const int id = s->token_ids[j];
        s->tok_str[j] = (id >= 0 && id < V) ? id2token[id] : "";
        s->text += s->tok_str[j];
      }
      for (int j = 0; j < cnt; ++j) s->tok_ptr[j] = s->tok_str[j].c_str();
      // U+2581 -> space, strip
      std::string txt;
      for (size_t p = 0; p < s->text.size();) {
        if (p + 2 < s->text.size() + 0 && (unsigned char)s->text[p] == 0xE2 && (unsigned char)s->text[p + 1] == 0x96 &&
            (unsigned char)s->text[p + 2] == 0x81) { txt += ' '; p += 3; }
        else txt += s->text[p++];
      }
      size_t a = txt.find_first_not_of(' '), b = txt.find_last_not_of(' ');
      s->text = (a == std::string::npos) ? "" : txt.substr(a, b - a + 1);
      s->json.clear();
      s->json_built = false;   // built on first request of the result (B200AsrGetOfflineStreamResult / ...AsJson)
      s->res.text = s->text.c_str(); s->res.json = nullptr; s->res.tokens = s->tok_ptr.data();
      s->res.token_ids = s->token_ids.data(); s->res.timestamps = s->timestamps.data(); s->res.frames = s->frames.data();
      s->res.ys_log_probs = s->lps.data(); s->res.tsallis = s->tsallis.data(); s->res.margin = s->margin.data();
      s->res.entropy = s->entropy.data(); s->res.top1 = s->top1.data(); s->res.count = cnt; s->res.num_frames = Tp[i];
      s->res.duration = dur;
      s->decoded = true;
    }
}

// Streams that carry their own hotwords (B200AsrCreateOfflineStreamWithHotwords) are decoded in one pass per distinct
// automaton; all others share the recognizer's.
void Engine::decode(Stream *const *ss, int n) {
  if (n <= 0) return;
  std::lock_guard<std::mutex> lk(mu);
  CUDA_CHECK(cudaSetDevice(device));
  bool mixed = false;
  for (int i = 0; i < n && !mixed; ++i) mixed = ss[i]->graph != nullptr || ss[i]->has_feats() != ss[0]->has_feats();
  if (!mixed) {
    pass_graph = default_graph();
    decode_part(ss, n);
    return;
  }
  std::vector<const GraphPair *> keys;
  std::vector<bool> kfeat;
  std::vector<std::vector<Stream *>> parts;
  for (int i = 0; i < n; ++i) {
    const GraphPair *k = ss[i]->graph.get();
    const bool f = ss[i]->has_feats();
    size_t j = 0;
    while (j < keys.size() && !(keys[j] == k && kfeat[j] == f)) ++j;
    if (j == keys.size()) { keys.push_back(k); kfeat.push_back(f); parts.emplace_back(); }
    parts[j].push_back(ss[i]);
  }
  for (size_t j = 0; j < parts.size(); ++j) {
    pass_graph = keys[j] ? (keys[j]->on_device ? &keys[j]->dev : nullptr) : default_graph();
    decode_part(parts[j].data(), (int)parts[j].size());
  }
  pass_graph = default_graph();
}

void Engine::decode_part(Stream *const *ss, int n) {
  // Batches bounded by total audio so workspaces stay bounded (about 70 min of audio per encoder pass). A call that needs
  // several of them runs up to kMaxLanes as ONE pass with forced groups: the search of batch k (on its lane's stream) overlaps
  // the fbank + encoder of batch k + 1.
  const long long kMaxSamples = 4200LL * 16000;
  static const bool chain = !(getenv("B200ASR_CHAIN_BATCHES") && atoi(getenv("B200ASR_CHAIN_BATCHES")) == 0);
  int begin = 0;
  while (begin < n) {
    std::vector<std::vector<int>> groups;
    int end = begin;
    const bool from_feats = ss[begin]->has_feats();    // a part is homogeneous (decode() partitions)
    while (end < n && (int)groups.size() < (chain && !from_feats ? kMaxLanes : 1)) {
      std::vector<int> g;
      long long tot = 0;
      while (end < n && (g.empty() || tot + ss[end]->n_samples() <= kMaxSamples)) {
        tot += ss[end]->n_samples();
        g.push_back(end - begin);
        ++end;
      }
      groups.push_back(std::move(g));
    }
    const int nb = end - begin;
    const auto hp0 = std::chrono::steady_clock::now();
    std::vector<long long> soff(nb + 1, 0), slen(nb, 0);
    for (int i = 0; i < nb; ++i) {
      slen[i] = ss[begin + i]->n_samples();
      soff[i + 1] = soff[i] + slen[i];
    }
    if (from_feats) {
      std::vector<long long> foff(nb + 1, 0);
      for (int i = 0; i < nb; ++i) {
        Stream *s = ss[begin + i];
        if ((slen[i] + 80) / 160 != s->feat_T) throw std::runtime_error("stream features do not match their sample count");
        foff[i + 1] = foff[i] + s->feat_T;
      }
      float *d_in = b_featin.get<float>((size_t)std::max<long long>(foff[nb], 1) * 80);
      CUDA_CHECK(cudaEventRecord(ev[4], st));
      for (int i = 0; i < nb; ++i)
        CUDA_CHECK(cudaMemcpyAsync(d_in + foff[i] * 80, ss[begin + i]->feats.data(), (size_t)ss[begin + i]->feat_T * 80 * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaEventRecord(ev[5], st));
      std::vector<int> Tp;
      soff.resize(nb);
      decode_pcm_device(nullptr, soff, slen, nb, &Tp, d_in, &foff);
      cudaEventElapsedTime(&tm.h2d, ev[4], ev[5]);
      unpack_results(ss + begin, nb, Tp);
      begin = end;
      continue;
    }
    // Streams whose PCM accept_waveform has already sent to this device are read where they sit (offsets relative to
    // the staging buffer's base, lengths explicit); only if some stream has no device copy is the whole batch packed
    // into the staging buffer the old way.
    bool resident = true;
    for (int i = 0; i < nb; ++i) {
      Stream *s = ss[begin + i];
      if (!s->samples.empty() && !(s->samples.on_device() && s->samples.device == device)) { resident = false; break; }
    }
    float *d_pcm = b_pcm.get<float>((size_t)std::max<long long>(resident ? 1 : soff[nb], 1));
    CUDA_CHECK(cudaEventRecord(ev[4], st));
    if (resident) {
      for (int i = 0; i < nb; ++i) {
        const float *dp = ss[begin + i]->samples.empty() ? d_pcm : ss[begin + i]->samples.d;
        soff[i] = (long long)((reinterpret_cast<intptr_t>(dp) - reinterpret_cast<intptr_t>(d_pcm)) / (intptr_t)sizeof(float));
      }
      // the pass must see the uploads accept_waveform queued on the device's copy stream
      CUDA_CHECK(cudaEventRecord(ev_copy, DevicePcmPool::get(device).stream()));
      CUDA_CHECK(cudaStreamWaitEvent(st, ev_copy, 0));
      for (int i = 0; i < nb; ++i) ss[begin + i]->samples.in_flight = false;   // this pass ends synchronised, after them
    } else {
      for (int i = 0; i < nb; ++i)
        if (!ss[begin + i]->samples.empty())
          CUDA_CHECK(cudaMemcpyAsync(d_pcm + soff[i], ss[begin + i]->samples.data(), ss[begin + i]->samples.size() * sizeof(float),
                                     cudaMemcpyHostToDevice, st));
    }
    CUDA_CHECK(cudaEventRecord(ev[5], st));
    std::vector<int> Tp;
    soff.resize(nb);
    const auto hp1 = std::chrono::steady_clock::now();
    decode_pcm_device(d_pcm, soff, slen, nb, &Tp, nullptr, nullptr, groups.size() > 1 ? &groups : nullptr);
    const auto hp2 = std::chrono::steady_clock::now();
    cudaEventElapsedTime(&tm.h2d, ev[4], ev[5]);
    unpack_results(ss + begin, nb, Tp);
    {
      const auto hp3 = std::chrono::steady_clock::now();
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      tm.d2h = (float)ms(hp2, hp3);   // host-side result unpacking
      static const bool host_prof = getenv("B200ASR_HOST_PROF") != nullptr;
      if (host_prof)
        fprintf(stderr, "[b200asr host prof] %d streams in %d batch(es): stage+enqueue H2D %.2f ms | device pass (wall) %.2f ms | unpack results %.2f ms\n",
                nb, (int)groups.size(), ms(hp0, hp1), ms(hp1, hp2), ms(hp2, hp3));
    }
    begin = end;
  }
}

}  // namespace b200asr

// ====================================================================== C ABI
using namespace b200asr;

struct B200AsrOfflineRecognizer { Engine eng; };
struct B200AsrOfflineStream { Stream s; };

#define API_TRY try {
#define API_CATCH(ret)                                    \
  } catch (const std::exception &e) {                     \
    g_last_error = e.what();                              \
    return ret;                                           \
  } catch (...) {                                         \
    g_last_error = "unknown error";                       \
    return ret;                                           \
  }

static Engine *E(const B200AsrOfflineRecognizer *r) {
  if (!r) throw std::runtime_error("null recognizer");
  return const_cast<Engine *>(&r->eng);
}

namespace b200asr {
void set_last_error(const std::string &msg) { g_last_error = msg; }   // for entry points living in other translation units
}

extern "C" {

const char *B200AsrGetLastError(void) { return g_last_error.c_str(); }
const char *B200AsrVersion(void) { return "b200asr 0.1.0 (sm_100a)"; }

const B200AsrOfflineRecognizer *B200AsrCreateOfflineRecognizer(const B200AsrOfflineRecognizerConfig *config) {
  B200AsrOfflineRecognizer *r = nullptr;
  try {
    if (!config) throw std::runtime_error("null config");
    r = new B200AsrOfflineRecognizer();
    r->eng.load(config);
    return r;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    delete r;
    return nullptr;
  }
}

void B200AsrDestroyOfflineRecognizer(const B200AsrOfflineRecognizer *r) { delete const_cast<B200AsrOfflineRecognizer *>(r); }

int32_t B200AsrOfflineRecognizerSetConfig(const B200AsrOfflineRecognizer *r, const B200AsrOfflineRecognizerConfig *c) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  const std::string dm = c->decoding_method ? c->decoding_method : "";
  if (!dm.empty()) {
    if (dm != "greedy_search" && dm != "modified_beam_search") throw std::runtime_error("unsupported decoding_method: " + dm);
    e->decoding_method = dm;
  }
  if (c->max_active_paths > 0) {
    if (c->max_active_paths > 16) throw std::runtime_error("max_active_paths > 16 is not built");
    e->max_active_paths = c->max_active_paths;
  }
  if (c->hotwords_score > 0) e->hotwords_score = c->hotwords_score;
  if (c->blank_penalty == c->blank_penalty) e->blank_penalty = c->blank_penalty;   // NaN = leave unchanged
  return 0;
  API_CATCH(-1)
}

int32_t B200AsrSetHotwordsTokenIds(const B200AsrOfflineRecognizer *r, const int32_t *tokens, const int32_t *offsets,
                                   const float *scores, int32_t n) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  e->set_graph(tokens, offsets, scores, n);
  return 0;
  API_CATCH(-1)
}

const B200AsrOfflineStream *B200AsrCreateOfflineStream(const B200AsrOfflineRecognizer *r) {
  try {
    auto *s = new B200AsrOfflineStream();
    s->s.eng = E(r);
    return s;
  } catch (const std::exception &e) { g_last_error = e.what(); return nullptr; }
}
void B200AsrDestroyOfflineStream(const B200AsrOfflineStream *s) { delete const_cast<B200AsrOfflineStream *>(s); }

int32_t B200AsrAcceptWaveformOffline(const B200AsrOfflineStream *s, int32_t sample_rate, const float *samples, int32_t n) {
  if (!s) { g_last_error = "accept_waveform: null stream"; return -1; }
  if (n < 0 || (n > 0 && !samples)) { g_last_error = "accept_waveform: null samples"; return -1; }
  if (sample_rate != 16000) { g_last_error = "accept_waveform: only 16000 Hz is supported (no resampler on the path)"; return -1; }
  if (n == 0) return 0;
  auto *ms = const_cast<B200AsrOfflineStream *>(s);
  if (ms->s.has_feats()) { g_last_error = "accept_waveform: the stream already holds precomputed features"; return -1; }
  try { ms->s.samples.append(samples, (size_t)n); } catch (const std::exception &e) { g_last_error = e.what(); return -1; }
  ms->s.samples.upload(ms->s.eng->device);
  ms->s.decoded = false;
  return 0;
}

int32_t B200AsrAcceptFeaturesOffline(const B200AsrOfflineStream *s, const float *feats, int32_t num_frames, int32_t feature_dim,
                                     int64_t num_samples) {
  if (!s || !feats || num_frames < 0) { g_last_error = "accept_features: null argument"; return -1; }
  if (feature_dim != 80) { g_last_error = "accept_features: only 80-bin fbank is supported (core/asr_engine.py:710)"; return -1; }
  if ((num_samples + 80) / 160 != num_frames) { g_last_error = "accept_features: num_frames must be (num_samples + 80) / 160"; return -1; }
  auto *ms = const_cast<B200AsrOfflineStream *>(s);
  if (!ms->s.samples.empty()) { g_last_error = "accept_features: the stream already holds samples"; return -1; }
  try { ms->s.feats.assign(feats, feats + (size_t)num_frames * 80); } catch (const std::exception &e) { g_last_error = e.what(); return -1; }
  ms->s.feat_T = num_frames;
  ms->s.feat_samples = num_samples;
  ms->s.decoded = false;
  return 0;
}

int32_t B200AsrAcceptWaveformsOffline(const B200AsrOfflineStream *const *ss, int32_t sample_rate, const float *const *samples,
                                      const int32_t *ns, int32_t n) {
  if (n < 0 || (n > 0 && (!ss || !samples || !ns))) { g_last_error = "accept_waveforms: null argument"; return -1; }
  for (int i = 0; i < n; ++i)
    if (B200AsrAcceptWaveformOffline(ss[i], sample_rate, samples[i], ns[i]) != 0) return -1;
  return 0;
}

// hotwords: phrases separated by '/', each a list of space-separated token ids with an optional " :score"
const B200AsrOfflineStream *B200AsrCreateOfflineStreamWithHotwords(const B200AsrOfflineRecognizer *r, const char *hotwords) {
  B200AsrOfflineStream *s = nullptr;
  try {
    Engine *e = E(r);
    s = new B200AsrOfflineStream();
    s->s.eng = e;
    const std::string hw = hotwords ? hotwords : "";
    std::vector<int32_t> toks, offs{0};
    std::vector<float> scs;
    std::stringstream all(hw);
    std::string phrase;
    while (std::getline(all, phrase, '/')) {
      float sc = e->hotwords_score;
      const size_t colon = phrase.rfind(':');
      if (colon != std::string::npos) { sc = (float)atof(phrase.c_str() + colon + 1); phrase = phrase.substr(0, colon); }
      std::stringstream ps(phrase);
      std::string tok;
      int cnt = 0;
      while (ps >> tok) {
        char *end = nullptr;
        const long id = strtol(tok.c_str(), &end, 10);
        if (end == tok.c_str() || *end != 0)
          throw std::runtime_error("stream hotwords must be token ids (\"12 34/56 78 :2.0\"); text phrases are tokenised by the host binding");
        toks.push_back((int32_t)id);
        ++cnt;
      }
      if (cnt == 0) continue;
      offs.push_back((int32_t)toks.size());
      scs.push_back(sc);
    }
    if (!scs.empty()) {
      std::lock_guard<std::mutex> lk(e->mu);
      CUDA_CHECK(cudaSetDevice(e->device));
      s->s.graph = e->make_graph(toks.data(), offs.data(), scs.data(), (int)scs.size());
    }
    return s;
  } catch (const std::exception &e) { g_last_error = e.what(); delete s; return nullptr; }
}

int32_t B200AsrDecodeMultipleOfflineStreams(const B200AsrOfflineRecognizer *r, const B200AsrOfflineStream *const *ss, int32_t n) {
  API_TRY
  Engine *e = E(r);
  std::vector<Stream *> v(n);
  for (int i = 0; i < n; ++i) {
    if (!ss[i]) throw std::runtime_error("null stream");
    v[i] = &const_cast<B200AsrOfflineStream *>(ss[i])->s;
  }
  e->decode(v.data(), n);
  return 0;
  API_CATCH(-1)
}
int32_t B200AsrDecodeOfflineStream(const B200AsrOfflineRecognizer *r, const B200AsrOfflineStream *s) {
  return B200AsrDecodeMultipleOfflineStreams(r, &s, 1);
}

const B200AsrOfflineRecognizerResult *B200AsrGetOfflineStreamResult(const B200AsrOfflineStream *s) {
  if (!s || !s->s.decoded) { g_last_error = "stream has not been decoded"; return nullptr; }
  build_json(const_cast<Stream *>(&s->s));
  return &s->s.res;
}
void B200AsrDestroyOfflineRecognizerResult(const B200AsrOfflineRecognizerResult *) {}
const char *B200AsrGetOfflineStreamResultAsJson(const B200AsrOfflineStream *s) {
  if (!s || !s->s.decoded) { g_last_error = "stream has not been decoded"; return nullptr; }
  build_json(const_cast<Stream *>(&s->s));
  return strdup(s->s.json.c_str());
}
void B200AsrDestroyOfflineStreamResultJson(const char *s) { free(const_cast<char *>(s)); }

int32_t B200AsrVocabSize(const B200AsrOfflineRecognizer *r) { return r ? r->eng.V : -1; }
int32_t B200AsrEncoderOutDim(const B200AsrOfflineRecognizer *r) { return r ? r->eng.join_dim : -1; }

// ---- raw stage entry points
int32_t B200AsrFbankBatch(const B200AsrOfflineRecognizer *r, const float *samples, const int64_t *sample_offsets, int32_t n,
                          float *out, int64_t *frame_offsets) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  std::vector<long long> soff(n + 1);
  for (int i = 0; i <= n; ++i) soff[i] = sample_offsets[i] - sample_offsets[0];
  long long total_frames = 0;
  std::vector<long long> foff(n + 1, 0);
  for (int i = 0; i < n; ++i) { foff[i + 1] = foff[i] + (soff[i + 1] - soff[i] + 80) / 160; }
  total_frames = foff[n];
  if (frame_offsets) for (int i = 0; i <= n; ++i) frame_offsets[i] = foff[i];
  if (!out) return (int32_t)total_frames;
  float *d_pcm = e->b_pcm.get<float>((size_t)std::max<long long>(soff[n], 1));
  long long *d_soff = e->b_soff.get<long long>(n + 1);
  CUDA_CHECK(cudaMemcpyAsync(d_pcm, samples + sample_offsets[0], soff[n] * sizeof(float), cudaMemcpyHostToDevice, e->st));
  CUDA_CHECK(cudaMemcpyAsync(d_soff, soff.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, e->st));
  float *d_feats = nullptr;
  std::vector<int> T;
  e->host_keep.clear();
  std::vector<long long> slen(n);
  for (int i = 0; i < n; ++i) slen[i] = soff[i + 1] - soff[i];
  e->run_fbank(d_pcm, d_soff, nullptr, slen, n, &d_feats, &T);
  CUDA_CHECK(cudaMemcpyAsync(out, d_feats, (size_t)total_frames * 80 * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  CUDA_CHECK(cudaStreamSynchronize(e->st));
  return (int32_t)total_frames;
  API_CATCH(-1)
}

int32_t B200AsrFbank(const B200AsrOfflineRecognizer *r, const float *samples, int32_t n, float *out) {
  const int64_t offs[2] = {0, n};
  return B200AsrFbankBatch(r, samples, offs, 1, out, nullptr);
}

int32_t B200AsrEncoder(const B200AsrOfflineRecognizer *r, const float *feats, const int32_t *x_lens, int32_t n, float *out,
                       int32_t *out_lens) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  std::vector<int> T(x_lens, x_lens + n);
  long long tot = 0, totp = 0;
  for (int i = 0; i < n; ++i) {
    tot += T[i];
    const int T1 = T[i] >= 9 ? (T[i] - 7) / 2 : 0;
    const int tp = (T1 + 1) / 2;
    if (out_lens) out_lens[i] = tp;
    totp += tp;
  }
  if (!out) return (int32_t)totp;
  float *d_feats = e->b_feats.get<float>((size_t)std::max<long long>(tot, 1) * 80);
  CUDA_CHECK(cudaMemcpyAsync(d_feats, feats, (size_t)tot * 80 * sizeof(float), cudaMemcpyHostToDevice, e->st));
  float *d_enc = nullptr;
  std::vector<int> Tp;
  e->gemm_flops = 0; e->gemm_launches = 0; e->gemm_ev_used = 0;
  e->host_keep.clear();
  e->pin_used = 0;
  e->reset_tile_counters();
  for (int attempt = 0; attempt < 2; ++attempt) {
    e->run_encoder(d_feats, T, e->b_enc, &d_enc, &Tp);
    CUDA_CHECK(cudaMemcpyAsync(out, d_enc, (size_t)totp * e->join_dim * sizeof(float), cudaMemcpyDeviceToHost, e->st));
    CUDA_CHECK(cudaStreamSynchronize(e->st));
    if (!e->softmax_overflowed()) break;     // else: once more with the exact two-pass softmax
    e->host_keep.clear();
    e->pin_used = 0;
    e->reset_tile_counters();
  }
  return (int32_t)totp;
  API_CATCH(-1)
}

int32_t B200AsrEncoderTap(const B200AsrOfflineRecognizer *r, const char *name, float *out, int32_t *dim) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  const std::string nm = name ? name : "";
  const int M1 = e->h_off[0].empty() ? 0 : e->h_off[0].back();
  const float *src = nullptr;
  int D = 0;
  if (nm == "embed") { src = e->b_x0.ptr<float>(); D = e->enc_dim[0]; }
  else if (nm.compare(0, 5, "stack") == 0) {
    const int i = atoi(nm.c_str() + 5);
    if (i < 0 || i >= (int)e->stacks.size()) throw std::runtime_error("no such tap: " + nm);
    src = e->b_stack[i].ptr<float>(); D = e->enc_dim[i];
  } else throw std::runtime_error("no such tap: " + nm);
  if (dim) *dim = D;
  if (out && M1 > 0) CUDA_CHECK(cudaMemcpy(out, src, (size_t)M1 * D * sizeof(float), cudaMemcpyDeviceToHost));
  return M1;
  API_CATCH(-1)
}

int32_t B200AsrDecoder(const B200AsrOfflineRecognizer *r, const int64_t *y, int32_t m, float *out) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  if (m <= 0) return 0;
  for (int i = 0; i < 2 * m; ++i) if (y[i] >= e->V) throw std::runtime_error("decoder: token id out of range");
  long long *dy = e->b_tmp.get<long long>((size_t)2 * m + (size_t)m * e->join_dim);
  float *dout = reinterpret_cast<float *>(dy + 2 * m);
  CUDA_CHECK(cudaMemcpyAsync(dy, y, (size_t)2 * m * sizeof(long long), cudaMemcpyHostToDevice, e->st));
  launch_decoder_rows(e->sm, dy, m, dout, e->st);
  CUDA_CHECK(cudaMemcpyAsync(out, dout, (size_t)m * e->join_dim * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  CUDA_CHECK(cudaStreamSynchronize(e->st));
  return 0;
  API_CATCH(-1)
}

int32_t B200AsrJoiner(const B200AsrOfflineRecognizer *r, const float *enc, const float *dec, int32_t m, float *logits) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  if (m <= 0) return 0;
  const size_t jd = e->join_dim;
  float *buf = e->b_tmp.get<float>((size_t)m * (3 * jd + e->V));
  float *d_enc = buf, *d_dec = buf + m * jd, *d_tmp = buf + 2 * m * jd, *d_lg = buf + 3 * m * jd;
  CUDA_CHECK(cudaMemcpyAsync(d_enc, enc, m * jd * sizeof(float), cudaMemcpyHostToDevice, e->st));
  CUDA_CHECK(cudaMemcpyAsync(d_dec, dec, m * jd * sizeof(float), cudaMemcpyHostToDevice, e->st));
  launch_joiner_rows(e->sm, d_enc, d_dec, m, d_tmp, d_lg, e->st);
  CUDA_CHECK(cudaMemcpyAsync(logits, d_lg, (size_t)m * e->V * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  CUDA_CHECK(cudaStreamSynchronize(e->st));
  return 0;
  API_CATCH(-1)
}

int32_t B200AsrDecoderJoinerInput(const B200AsrOfflineRecognizer *r, const int64_t *y, const float *enc, int32_t m, float *dec_out,
                                  float *x_out) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  if (m <= 0) return 0;
  for (int i = 0; i < 2 * m; ++i) if (y[i] >= e->V) throw std::runtime_error("decoder: token id out of range");
  const size_t jd = e->join_dim;
  long long *dy = e->b_tmp.get<long long>((size_t)2 * m + 3 * (size_t)m * jd);
  float *d_enc = reinterpret_cast<float *>(dy + 2 * m), *d_dec = d_enc + m * jd, *d_x = d_dec + m * jd;
  CUDA_CHECK(cudaMemcpyAsync(dy, y, (size_t)2 * m * sizeof(long long), cudaMemcpyHostToDevice, e->st));
  if (enc) CUDA_CHECK(cudaMemcpyAsync(d_enc, enc, m * jd * sizeof(float), cudaMemcpyHostToDevice, e->st));
  e->ensure_dec_table();
  launch_decoder_product_rows(e->sm, dy, enc ? d_enc : nullptr, m, d_dec, d_x, e->st);
  if (dec_out) CUDA_CHECK(cudaMemcpyAsync(dec_out, d_dec, m * jd * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  if (x_out) CUDA_CHECK(cudaMemcpyAsync(x_out, d_x, m * jd * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  CUDA_CHECK(cudaStreamSynchronize(e->st));
  return 0;
  API_CATCH(-1)
}

int32_t B200AsrJoinerRecords(const B200AsrOfflineRecognizer *r, const float *x, int32_t m, int32_t kb, float *records) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  if (kb != 4 && kb != 8 && kb != 16) throw std::runtime_error("kb must be 4, 8 or 16");
  const int P = (e->V + kPartCols - 1) / kPartCols, REC = part_rec_floats(kb);
  if (m <= 0) return P * REC;
  if (!records) return P * REC;
  const size_t jd = e->join_dim;
  float *buf = e->b_tmp.get<float>((size_t)m * jd + (size_t)m * P * REC + 16);
  float *d_x = buf, *d_rec = buf + (((size_t)m * jd + 3) & ~size_t(3));
  CUDA_CHECK(cudaMemcpyAsync(d_x, x, m * jd * sizeof(float), cudaMemcpyHostToDevice, e->st));
  launch_joiner_records(e->search, e->sm, d_x, m, kb, d_rec, e->st);
  CUDA_CHECK(cudaMemcpyAsync(records, d_rec, (size_t)m * P * REC * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  CUDA_CHECK(cudaStreamSynchronize(e->st));
  return P * REC;
  API_CATCH(-1)
}

int32_t B200AsrBeamSearch(const B200AsrOfflineRecognizer *r, const float *enc_out, const int32_t *lens, int32_t n, int32_t method,
                          int32_t beam, int32_t max_tokens, int32_t *tokens, int32_t *frames, float *tok_logprobs, float *stats,
                          int32_t *n_tokens) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  long long tot = 0;
  for (int i = 0; i < n; ++i) tot += std::max(lens[i], 0);
  float *d_enc = e->b_enc.get<float>((size_t)std::max<long long>(tot, 1) * e->join_dim);
  CUDA_CHECK(cudaMemcpyAsync(d_enc, enc_out, (size_t)tot * e->join_dim * sizeof(float), cudaMemcpyHostToDevice, e->st));
  SearchResultHost res{};
  res.n_utts = n; res.max_tokens = max_tokens; res.n_tokens = n_tokens; res.tokens = tokens; res.frames = frames;
  res.tok_lp = tok_logprobs; res.stats = stats;
  std::vector<int> l(lens, lens + n);
  e->ensure_dec_table();
  run_search(e->search, e->sm, e->default_graph(), d_enc, l.data(), n, method, beam, e->blank_penalty, &res, e->st);
  return 0;
  API_CATCH(-1)
}

int32_t B200AsrGemm(const B200AsrOfflineRecognizer *r, const float *A, const float *W, const float *bias, const float *R, float *C,
                    int32_t M, int32_t N, int32_t K, int32_t act, int32_t impl, int32_t reps, float *ms_per_launch) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  const size_t nA = (size_t)M * K, nW = (size_t)N * K, nC = (size_t)M * N;
  float *buf = e->b_tmp.get<float>(nA + nW + 2 * nC + N + 64);
  float *dA = buf, *dW = dA + ((nA + 3) & ~size_t(3)), *dC = dW + ((nW + 3) & ~size_t(3)), *dR = dC + ((nC + 3) & ~size_t(3));
  float *dB = dR + ((nC + 3) & ~size_t(3));
  CUDA_CHECK(cudaMemcpyAsync(dA, A, nA * 4, cudaMemcpyHostToDevice, e->st));
  CUDA_CHECK(cudaMemcpyAsync(dW, W, nW * 4, cudaMemcpyHostToDevice, e->st));
  if (bias) CUDA_CHECK(cudaMemcpyAsync(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice, e->st));
  if (R) CUDA_CHECK(cudaMemcpyAsync(dR, R, nC * 4, cudaMemcpyHostToDevice, e->st));
  GemmArgs g{};
  g.A = dA; g.lda = K; g.W = dW; g.bias = bias ? dB : nullptr; g.R = R ? dR : nullptr; g.ldr = N; g.C = dC; g.ldc = N;
  g.M = M; g.N = N; g.K = K; g.act = act;
  if (impl == 2 || impl == 3) {
    float *dWlo = e->b_tmp2.get<float>(nW);
    launch_split_lo(dW, dWlo, (long long)nW, e->st);
    g.Wlo = dWlo;
  }
  void *w16hi = nullptr, *w16lo = nullptr;
  if (impl == 3 || impl == 4) {
    int ld = 0;
    split_weights_16(dW, N, K, impl == 4, &w16hi, &w16lo, &ld, e->st);
    g.W16hi = w16hi; g.W16lo = w16lo; g.w16_ld = ld;
  }
  if (reps < 1) reps = 1;
  auto run = [&]() {
    if (impl == 1) launch_gemm_tc(g, e->st);
    else if (impl == 2 || impl == 3) launch_gemm_tc3(g, e->st);      // 3 carries the 16-bit copies -> fp16 operand split
    else if (impl == 4) launch_gemm_bf16(g, e->st);
    else launch_gemm_fp32(g, e->st);
  };
  run();   // warm-up / the checked result
  CUDA_CHECK(cudaEventRecord(e->ev[6], e->st));
  for (int i = 1; i < reps; ++i) run();
  CUDA_CHECK(cudaEventRecord(e->ev[7], e->st));
  CUDA_CHECK(cudaMemcpyAsync(C, dC, nC * 4, cudaMemcpyDeviceToHost, e->st));
  CUDA_CHECK(cudaStreamSynchronize(e->st));
  cudaFree(w16hi); cudaFree(w16lo);
  if (ms_per_launch) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev[6], e->ev[7]);
    *ms_per_launch = reps > 1 ? ms / (reps - 1) : 0.f;
  }
  return 0;
  API_CATCH(-1)
}

double B200AsrContextForwardOneStep(const B200AsrOfflineRecognizer *r, int32_t state, int32_t token, int32_t *next_state) {
  if (!r || r->eng.cg_host().n_nodes() == 0 || state < 0 || state >= r->eng.cg_host().n_nodes()) { if (next_state) *next_state = 0; return 0.0; }
  int nxt = 0;
  const double d = cg_forward_one_step(r->eng.cg_host().view(), state, token, &nxt);
  if (next_state) *next_state = nxt;
  return d;
}
double B200AsrContextFinalize(const B200AsrOfflineRecognizer *r, int32_t state) {
  if (!r || r->eng.cg_host().n_nodes() == 0 || state < 0 || state >= r->eng.cg_host().n_nodes()) return 0.0;
  return cg_finalize(r->eng.cg_host().view(), state);
}
int32_t B200AsrContextNumNodes(const B200AsrOfflineRecognizer *r) { return r ? r->eng.cg_host().n_nodes() : 0; }

// ---- device-resident benchmarking hooks
int32_t B200AsrStageBatch(const B200AsrOfflineRecognizer *r, const float *samples, const int64_t *sample_offsets, int32_t n) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  Engine::Staged sgd;
  sgd.n = n;
  sgd.h_soff.resize(n + 1);
  for (int i = 0; i <= n; ++i) sgd.h_soff[i] = sample_offsets[i] - sample_offsets[0];
  CUDA_CHECK(cudaMalloc(&sgd.pcm, std::max<long long>(sgd.h_soff[n], 1) * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&sgd.soff, (n + 1) * sizeof(long long)));
  CUDA_CHECK(cudaMemcpy(sgd.pcm, samples + sample_offsets[0], sgd.h_soff[n] * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(sgd.soff, sgd.h_soff.data(), (n + 1) * sizeof(long long), cudaMemcpyHostToDevice));
  const int h = e->next_handle++;
  e->staged[h] = std::move(sgd);
  return h;
  API_CATCH(-1)
}

int32_t B200AsrRunStagedBatch(const B200AsrOfflineRecognizer *r, int32_t handle, int32_t *n_tokens) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  auto it = e->staged.find(handle);
  if (it == e->staged.end()) throw std::runtime_error("no such staged batch");
  const Engine::Staged &sgd = it->second;
  std::vector<int> Tp;
  std::vector<long long> slen(sgd.n), soff(sgd.h_soff.begin(), sgd.h_soff.begin() + sgd.n);
  for (int i = 0; i < sgd.n; ++i) slen[i] = sgd.h_soff[i + 1] - sgd.h_soff[i];
  e->pass_graph = e->default_graph();
  e->decode_pcm_device(sgd.pcm, soff, slen, sgd.n, &Tp);
  if (n_tokens) for (int i = 0; i < sgd.n; ++i) n_tokens[i] = e->result_of(i).n_tokens;
  return 0;
  API_CATCH(-1)
}

/* `reps` passes over a staged batch issued back to back, the search of pass k (on its own stream) beside the encoder of pass
 * k + 1; returns after the last search. total_ms = device time from the first fbank launch to the end of the last search.
 * n_tokens (may be NULL) receives the token counts of the last pass. reps <= 8. */
int32_t B200AsrRunStagedBatchChained(const B200AsrOfflineRecognizer *r, int32_t handle, int32_t reps, int32_t *n_tokens, float *total_ms) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  CUDA_CHECK(cudaSetDevice(e->device));
  auto it = e->staged.find(handle);
  if (it == e->staged.end()) throw std::runtime_error("no such staged batch");
  if (reps < 1 || reps > kMaxLanes) throw std::runtime_error("reps must be in [1, 8]");
  const Engine::Staged &sgd = it->second;
  const int n = sgd.n;
  std::vector<long long> slen((size_t)n * reps), soff((size_t)n * reps);
  std::vector<std::vector<int>> groups(reps);
  for (int k = 0; k < reps; ++k)
    for (int i = 0; i < n; ++i) {
      soff[(size_t)k * n + i] = sgd.h_soff[i];
      slen[(size_t)k * n + i] = sgd.h_soff[i + 1] - sgd.h_soff[i];
      groups[k].push_back(k * n + i);
    }
  std::vector<int> Tp;
  e->pass_graph = e->default_graph();
  e->decode_pcm_device(sgd.pcm, soff, slen, n * reps, &Tp, nullptr, nullptr, reps > 1 ? &groups : nullptr);
  if (n_tokens) for (int i = 0; i < n; ++i) n_tokens[i] = e->result_of((reps - 1) * n + i).n_tokens;
  if (total_ms) *total_ms = e->tm.total;
  return 0;
  API_CATCH(-1)
}

/* Token ids of utterance u of the last staged run / decode pass (bench.py's parity self-check). Returns the count. */
int32_t B200AsrLastPassTokens(const B200AsrOfflineRecognizer *r, int32_t u, int32_t *tokens, int32_t *frames, int32_t cap) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  if (u < 0 || u >= (int)e->lane_of.size()) throw std::runtime_error("no such utterance in the last pass");
  const Engine::UttResult res = e->result_of(u);
  for (int j = 0; j < res.n_tokens && j < cap; ++j) {
    if (tokens) tokens[j] = res.tokens[j];
    if (frames) frames[j] = res.frames[j];
  }
  return res.n_tokens;
  API_CATCH(-1)
}

/* Pipeline shape of the last pass: number of groups, per-group search time (ms, on its own stream), their sum, and the
 * device->host result bytes. */
int32_t B200AsrLastPipelineStats(const B200AsrOfflineRecognizer *r, int32_t *n_groups, float *search_busy_ms, float *lane_ms8,
                                 int64_t *d2h_bytes) {
  if (!r) return -1;
  const Engine &e = r->eng;
  if (n_groups) *n_groups = e.n_groups_last;
  if (search_busy_ms) *search_busy_ms = e.search_busy_ms;
  if (lane_ms8) for (int i = 0; i < kMaxLanes; ++i) lane_ms8[i] = i < e.n_groups_last ? e.lane_ms[i] : 0.f;
  if (d2h_bytes) *d2h_bytes = e.d2h_bytes_last;
  return 0;
}

int32_t B200AsrLastPipelineTimeline(const B200AsrOfflineRecognizer *r, float *t_ms, int32_t max_groups) {
  if (!r || !t_ms) return -1;
  const Engine &e = r->eng;
  const int n = std::min<int>(e.n_groups_last, max_groups);
  for (int g = 0; g < n; ++g) for (int k = 0; k < 4; ++k) t_ms[g * 4 + k] = e.lane_t[g][k];
  return n;
}

int32_t B200AsrReleaseBatch(const B200AsrOfflineRecognizer *r, int32_t handle) {
  API_TRY
  Engine *e = E(r);
  std::lock_guard<std::mutex> lk(e->mu);
  auto it = e->staged.find(handle);
  if (it == e->staged.end()) return -1;
  cudaFree(it->second.pcm); cudaFree(it->second.soff);
  e->staged.erase(it);
  return 0;
  API_CATCH(-1)
}

int32_t B200AsrLastTimings(const B200AsrOfflineRecognizer *r, float *out6, int64_t *n_launches) {
  if (!r) return -1;
  const Engine &e = r->eng;
  if (out6) { out6[0] = e.tm.fbank; out6[1] = e.tm.encoder; out6[2] = e.tm.search; out6[3] = e.tm.total; out6[4] = e.tm.h2d; out6[5] = e.tm.d2h; }
  if (n_launches) *n_launches = e.launches_last;
  return 0;
}
int32_t B200AsrLastGemmStats(const B200AsrOfflineRecognizer *r, double *ms, double *flops, int64_t *launches) {
  if (!r) return -1;
  if (ms) *ms = r->eng.gemm_ms;
  if (flops) *flops = r->eng.gemm_flops;
  if (launches) *launches = r->eng.gemm_launches;
  return 0;
}
double B200AsrLastGemmBytes(const B200AsrOfflineRecognizer *r) { return r ? r->eng.gemm_bytes : -1.0; }
int32_t B200AsrSetProfiling(const B200AsrOfflineRecognizer *r, int32_t on) {
  if (!r) return -1;
  const_cast<Engine &>(r->eng).profiling = on != 0;
  return 0;
}

}  // extern "C"
