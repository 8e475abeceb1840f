// Host side of batch_obs (reference: ss_baselines/common/utils.py:129-156): the per-env observation arrays a VectorEnv
// hands the trainer are gathered into ONE pinned staging buffer per sensor by a small persistent pool of native
// threads.  At 64 envs a step's frames are 7.3 MB in 128 pieces; copied piece by piece from Python (numpy assignment,
// also from a Python thread pool: the interpreter overhead per piece and the GIL hand-offs dominate) the staging took
// longer than the whole policy step takes the GPU.  Disjoint byte ranges per thread, streaming stores (see copy_stream), no CUDA calls.
#include "common.cuh"

#ifndef AVL_HOST_EMUL
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace {

// Copy with non-temporal stores.  The destination is pinned staging memory that the GPU's copy engine reads next: with
// ordinary stores the freshly written lines sit dirty in the cores' caches and the DMA read has to snoop them out
// (measured on the GPU box: the 7.3 MB H2D copy of a step ran at 8-12 GB/s instead of 53 GB/s); streaming stores go
// to DRAM and leave the caches alone.  Pieces are 48-64 KB — far below the size at which memcpy switches by itself.
inline void copy_stream(char* dst, const char* src, size_t n) {
#if defined(__x86_64__)
  if (n < 256) {
    memcpy(dst, src, n);
    return;
  }
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
  if (head) {
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
  }
  const size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
    src += 64; dst += 64;
  }
  const size_t tail = n - blocks * 64;
  if (tail) memcpy(dst, src, tail);
#else
  memcpy(dst, src, n);
#endif
}
inline void copy_fence() {
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

static int g_stream_stores = 1;

struct GatherJob {
  const void* const* src;
  void* const* dst;
  const long long* bytes;
  int n;
};

class Pool {
 public:
  explicit Pool(int workers) : stop_(false), job_{}, generation_(0), pending_(0), workers_n_(workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int workers() const { return workers_n_; }
  // the caller takes share `workers` itself, so workers + 1 threads copy
  void run(const GatherJob& j) {
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = j;
      pending_ = workers_n_;
      ++generation_;
    }
    cv_.notify_all();
    copy_share(j, workers_n_, workers_n_ + 1);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  // share k of `parts`: a contiguous range of the concatenated byte stream, cut at 4 KB granules
  static void copy_share(const GatherJob& j, int k, int parts) {
    long long total = 0;
    for (int i = 0; i < j.n; ++i) total += j.bytes[i];
    const long long granule = 4096;
    const long long per = ((total + parts - 1) / parts + granule - 1) / granule * granule;
    long long lo = per * k, hi = lo + per;
    if (hi > total) hi = total;
    long long pos = 0;
    for (int i = 0; i < j.n && pos < hi; ++i) {
      const long long b = j.bytes[i];
      const long long a0 = lo > pos ? lo - pos : 0;
      const long long a1 = (hi - pos) < b ? (hi - pos) : b;
      if (a1 > a0) {
        if (g_stream_stores) copy_stream(static_cast<char*>(j.dst[i]) + a0, static_cast<const char*>(j.src[i]) + a0, (size_t)(a1 - a0));
        else memcpy(static_cast<char*>(j.dst[i]) + a0, static_cast<const char*>(j.src[i]) + a0, (size_t)(a1 - a0));
      }
      pos += b;
    }
    copy_fence();
  }
  void loop(int k) {
    long long seen = 0;
    for (;;) {
      GatherJob j;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        j = job_;
      }
      copy_share(j, k, workers_n_ + 1);
      {
        std::lock_guard<std::mutex> lk(m_);
        --pending_;
      }
      done_.notify_one();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_;
  bool stop_;
  GatherJob job_;
  long long generation_;
  int pending_;
  int workers_n_;
  std::vector<std::thread> threads_;
};

Pool* g_pool = nullptr;
std::mutex g_pool_mutex;  // one gather at a time (the pool holds a single job slot)
int g_pool_threads = 0;   // 0: automatic

}  // namespace

// Copy threads of avl_host_gather (the caller included): 0 = automatic (hardware threads / 2, at most 8).  Takes effect
// at the next gather.  Returns the old setting.
AVL_API int avl_set_host_gather_threads(int threads) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  int old = g_pool_threads;
  g_pool_threads = threads < 0 ? 0 : threads;
  if (g_pool) {
    delete g_pool;
    g_pool = nullptr;
  }
  return old;
}

// 1 (default): the gather writes with non-temporal stores (the destination is read next by the GPU's copy engine, not by
// the CPU); 0: plain memcpy.  Returns the old setting.
AVL_API int avl_set_host_gather_streaming(int on) {
  int old = g_stream_stores;
  g_stream_stores = on ? 1 : 0;
  return old;
}

// dst[i][0 .. bytes[i]) = src[i][0 .. bytes[i]) for n pieces of host memory, byte ranges split evenly over the copy
// threads.  Pieces must not overlap.  No CUDA call is made; returns once every byte is copied.
AVL_API int avl_host_gather(const void* const* src, void* const* dst, const long long* bytes, int n) {
  if (n < 0) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!src || !dst || !bytes) return AVL_ERR_ARG;
  long long total = 0;
  for (int i = 0; i < n; ++i) {
    if (bytes[i] < 0 || (bytes[i] > 0 && (!src[i] || !dst[i]))) return AVL_ERR_ARG;
    total += bytes[i];
  }
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  int threads = g_pool_threads;
  if (threads == 0) {
    unsigned hw = std::thread::hardware_concurrency();
    threads = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
  }
  if (threads <= 1 || total < (1 << 18)) {  // small: the calling thread alone
    const bool stream = g_stream_stores && total >= (1 << 18);
    for (int i = 0; i < n; ++i) {
      if (bytes[i] <= 0) continue;
      if (stream) copy_stream(static_cast<char*>(dst[i]), static_cast<const char*>(src[i]), (size_t)bytes[i]);
      else memcpy(dst[i], src[i], (size_t)bytes[i]);
    }
    copy_fence();
    return AVL_OK;
  }
  if (!g_pool || g_pool->workers() != threads - 1) {
    delete g_pool;
    g_pool = new Pool(threads - 1);
  }
  GatherJob j{src, dst, bytes, n};
  g_pool->run(j);
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
