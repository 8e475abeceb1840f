// Library-wide state: last CUDA error, device properties, version.
#include "common.cuh"

static int g_last_cuda_error = 0;
static int g_num_sms = 0;
static long long g_launches = 0;
static int g_use_tc = 1;

static long long g_config_epoch = 0;

extern "C" void avl_count_launch() { ++g_launches; }
extern "C" void avl_add_launches(long long n) { g_launches += n; }
long long avl_launch_count_internal() { return g_launches; }
// every avl_set_* toggle that changes WHICH kernels a call launches bumps the epoch: cached CUDA graphs of whole-network
// calls (resnet_fwd.cu) are keyed by it
extern "C" void avl_bump_config_epoch() { ++g_config_epoch; }
extern "C" long long avl_config_epoch() { return g_config_epoch; }

extern "C" int avl_set_cuda_error(int e) {
  g_last_cuda_error = e;
  return e;
}

int avl_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_num_sms = n;
  }
  return g_num_sms;
}

AVL_API int avl_version(void) { return 100; }

// cudaError_t of the last failed runtime call made by this library (0 = none).
AVL_API int avl_last_cuda_error(void) { return g_last_cuda_error; }

AVL_API const char* avl_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)g_last_cuda_error); }

AVL_API int avl_device_sm_count(void) { return avl_num_sms(); }

// Number of kernels this library has launched so far in this process (bench.py's gpu_launches).
AVL_API long long avl_launch_count(void) { return g_launches; }
// A caller that replays kernels of this library through a CUDA graph of its own (the trainer's whole-step graphs)
// reports the launches of each replay here, so that the count stays the number of kernels that actually ran.
AVL_API long long avl_launch_count_add(long long n) {
  g_launches += n;
  return g_launches;
}

// Tensor-core (tcgen05, TF32 operands) level: 0 = fp32 SIMT kernels only; 1 (default) = encoder convolutions /
// fully-connected layers (stated tolerance 2e-3 of the output range; the reference's cuDNN convolutions also run
// TF32 by default); 2 = additionally the dense layers of the scene-memory transformer (opt-in: fp32 outputs of
// that block are otherwise held to 1e-3).
AVL_API int avl_set_tensor_cores(int level) {
  int old = g_use_tc;
  g_use_tc = level < 0 ? 0 : (level > 2 ? 2 : level);
  avl_bump_config_epoch();
  return old;
}
AVL_API int avl_get_tensor_cores(void) { return g_use_tc; }

// Programmatic dependent launch between the kernels of the encoder chains (convolutions, GroupNorm): 1 (default) on, 0 off.
// Returns the old value.
static int g_pdl = 1;
extern "C" int avl_pdl_enabled() { return g_pdl; }
AVL_API int avl_set_pdl(int on) {
  int old = g_pdl;
  g_pdl = on ? 1 : 0;
  avl_bump_config_epoch();
  return old;
}

// 32-byte global stores (sm_100 STG.256) in the row-per-thread TMEM epilogues: 1 (default) on, 0 = 16-byte stores
// (diagnostic).  Returns the old value.
static int g_wide_stores = 1;
extern "C" int avl_wide_stores() { return g_wide_stores; }
AVL_API int avl_set_wide_stores(int on) {
  int old = g_wide_stores;
  g_wide_stores = on ? 1 : 0;
  avl_bump_config_epoch();
  return old;
}
