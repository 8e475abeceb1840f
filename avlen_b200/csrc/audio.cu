// Batched audiogoal rendering + log-magnitude spectrogram.
//
// Replaces, for a whole batch of environments in one launch:
//   soundspaces/simulator.py:644-699   SoundSpacesSim._compute_audiogoal
//   soundspaces/tasks/nav.py:87-101    SpectrogramSensor.compute_spectrogram
//
// One persistent CTA per SM walks over environments.  Per environment the
// binaural waveform is the causal FIR  y[c][n] = sum_k rir[k][c] * src[base+n-k]
// (all three reference branches reduce to this), evaluated
// as a 32768-point circular convolution through a 16384-point complex FFT that
// lives entirely in shared memory (packed-real trick, in-place DIF forward /
// DIT inverse so no digit-reversal pass is needed).  The (2, sr) waveform never
// leaves shared memory unless the audiogoal output is requested: the STFT
// (n_fft 512, hop 160, periodic Hann 400 centred, reflect padding), |.|, the
// zero-padded 4x4 block mean and log1p are computed from the shared-memory
// copy and only the (65, 26, 2) spectrogram is written.
#include "common.cuh"

namespace {

constexpr int kM = 16384;  // complex FFT length
constexpr int kP = 32768;  // circular convolution length (real samples)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBufElems = kM + (kM >> 4) + (kM >> 8);  // padded float2 count
constexpr int kFrameElems = 272;                        // 256 + 16 pad
constexpr int kNfft = 512, kHop = 160, kWin = 400, kBins = 257, kFB = 65;

typedef float2 cf;

__device__ __forceinline__ int padi(int p) { return p + (p >> 4) + (p >> 8); }
__device__ __forceinline__ int padf(int p) { return p + (p >> 4); }
__device__ __forceinline__ cf cmul(cf a, cf b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cf cmulc(cf a, cf b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ cf cadd(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cf csub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cf cconj(cf a) { return make_float2(a.x, -a.y); }

// y_q = sum_m x_m exp(SIGN * 2*pi*i*m*q/4), in place, natural order
template <int SIGN>
__device__ __forceinline__ void fft4(cf& a, cf& b, cf& c, cf& d) {
  cf t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  cf r3 = (SIGN < 0) ? make_float2(t3.y, -t3.x) : make_float2(-t3.y, t3.x);  // (+-i) * t3
  a = cadd(t0, t2);
  c = csub(t0, t2);
  b = cadd(t1, r3);
  d = csub(t1, r3);
}

// constant exp(SIGN*2*pi*i*k/16)
template <int SIGN>
__device__ __forceinline__ cf w16(int k) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
  float re, im;
  switch (k) {
    case 1: re = c1; im = s1; break;
    case 2: re = r; im = r; break;
    case 3: re = s1; im = c1; break;
    case 4: re = 0.f; im = 1.f; break;
    case 6: re = -r; im = r; break;
    case 9: re = -c1; im = -s1; break;
    default: re = 1.f; im = 0.f; break;
  }
  return make_float2(re, SIGN < 0 ? -im : im);
}

// 16-point DFT: y_q = sum_m x_m exp(SIGN*2*pi*i*m*q/16); y_q is left in x[o16(q)]
__host__ __device__ constexpr int o16(int q) { return 4 * (q & 3) + (q >> 2); }

template <int SIGN>
__device__ __forceinline__ void fft16(cf (&x)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    fft4<SIGN>(x[j], x[j + 4], x[j + 8], x[j + 12]);  // x[j + 4*q0] = t_{q0}[j]
#pragma unroll
    for (int q0 = 1; q0 < 4; ++q0)
      if (j > 0) x[j + 4 * q0] = cmul(x[j + 4 * q0], w16<SIGN>(j * q0));
  }
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) fft4<SIGN>(x[4 * q0], x[4 * q0 + 1], x[4 * q0 + 2], x[4 * q0 + 3]);
}

// tw[q] = w^q, q = 1..15 (multiplication tree, depth <= 4)
__device__ __forceinline__ void tw_powers(cf w, cf (&t)[16]) {
  t[0] = make_float2(1.f, 0.f);
  t[1] = w;
  t[2] = cmul(w, w);
  t[3] = cmul(t[2], w);
  t[4] = cmul(t[2], t[2]);
  t[5] = cmul(t[4], w);
  t[6] = cmul(t[3], t[3]);
  t[7] = cmul(t[4], t[3]);
  t[8] = cmul(t[4], t[4]);
  t[9] = cmul(t[8], w);
  t[10] = cmul(t[5], t[5]);
  t[11] = cmul(t[8], t[3]);
  t[12] = cmul(t[6], t[6]);
  t[13] = cmul(t[8], t[5]);
  t[14] = cmul(t[7], t[7]);
  t[15] = cmul(t[8], t[7]);
}

// One radix-16 butterfly of an in-place DIF stage on a block of size n = 16*s:
//   buf[first + q*s] <- (sum_m buf[first + m*s] w16^{mq}) * w^q
// addressed through the PADDED index of its first point and a compile-time padded stride SP: for every stage of the two
// transforms the padding functions are linear along a butterfly's 16 points (the stride is a multiple of the padding
// period, or the points stay inside one period), so padi / padf is evaluated once per butterfly instead of per point —
// the per-point shifts and adds were a third of the kernel's instructions (profiles/r02_audio_render_ncu_full.txt).
template <int SP>
__device__ __forceinline__ void r16_fwd(cf* buf, int p0, bool twiddle, cf w) {
  cf x[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) x[m] = buf[p0 + m * SP];
  fft16<-1>(x);
  if (twiddle) {
    cf t[16];
    tw_powers(w, t);
#pragma unroll
    for (int q = 1; q < 16; ++q) x[o16(q)] = cmul(x[o16(q)], t[q]);
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[p0 + q * SP] = x[o16(q)];
}

// Inverse of r16_fwd up to the factor 16.
template <int SP>
__device__ __forceinline__ void r16_inv(cf* buf, int p0, bool twiddle, cf w) {
  cf x[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) x[q] = buf[p0 + q * SP];
  if (twiddle) {
    cf t[16];
    tw_powers(w, t);
#pragma unroll
    for (int q = 1; q < 16; ++q) x[q] = cmulc(x[q], t[q]);
  }
  fft16<+1>(x);
#pragma unroll
  for (int m = 0; m < 16; ++m) buf[p0 + m * SP] = x[o16(m)];
}

// padded strides of the big buffer (padi(p) = p + p/16 + p/256): 4096 -> 4368, 256 -> 273, 16 -> 17 (inside one block of
// 256), 1 -> 1 (inside one block of 16); of a frame buffer (padf(p) = p + p/16): 16 -> 17, 1 -> 1
constexpr int kSP4096 = 4096 + 256 + 16, kSP256 = 256 + 16 + 1, kSP16 = 17;

// tw[k] = exp(-2*pi*i*k/32768), k in [0, 16384]
// In-place DIF forward FFT of the 16384-point buffer; radices 4,16,16,16.
// Output X[k], k = q1 + 4 q2 + 64 q3 + 1024 q4, is left at rev(k) = q1*4096 + q2*256 + q3*16 + q4.
__device__ void fft_big_fwd(cf* buf, const cf* __restrict__ tw) {
  const int tid = threadIdx.x;
  for (int u = tid; u < 4096; u += kThreads) {
    const int p = padi(u);
    cf a = buf[p], b = buf[p + kSP4096], c = buf[p + 2 * kSP4096], d = buf[p + 3 * kSP4096];
    fft4<-1>(a, b, c, d);
    cf w1 = tw[2 * u];
    cf w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    buf[p] = a;
    buf[p + kSP4096] = cmul(b, w1);
    buf[p + 2 * kSP4096] = cmul(c, w2);
    buf[p + 3 * kSP4096] = cmul(d, w3);
  }
  __syncthreads();
  for (int u = tid; u < 1024; u += kThreads) {
    int b = u >> 8, j = u & 255;
    r16_fwd<kSP256>(buf, padi(b * 4096 + j), true, tw[8 * j]);  // w_4096^j
  }
  __syncthreads();
  for (int u = tid; u < 1024; u += kThreads) {
    int b = u >> 4, j = u & 15;
    r16_fwd<kSP16>(buf, padi(b * 256 + j), true, tw[128 * j]);  // w_256^j
  }
  __syncthreads();
  for (int u = tid; u < 1024; u += kThreads) r16_fwd<1>(buf, padi(u * 16), false, make_float2(1.f, 0.f));
  __syncthreads();
}

// Exact inverse of fft_big_fwd up to the factor 16384 (digit-reversed in, natural out).
__device__ void fft_big_inv(cf* buf, const cf* __restrict__ tw) {
  const int tid = threadIdx.x;
  for (int u = tid; u < 1024; u += kThreads) r16_inv<1>(buf, padi(u * 16), false, make_float2(1.f, 0.f));
  __syncthreads();
  for (int u = tid; u < 1024; u += kThreads) {
    int b = u >> 4, j = u & 15;
    r16_inv<kSP16>(buf, padi(b * 256 + j), true, tw[128 * j]);
  }
  __syncthreads();
  for (int u = tid; u < 1024; u += kThreads) {
    int b = u >> 8, j = u & 255;
    r16_inv<kSP256>(buf, padi(b * 4096 + j), true, tw[8 * j]);
  }
  __syncthreads();
  for (int u = tid; u < 4096; u += kThreads) {
    const int p = padi(u);
    cf w1 = tw[2 * u];
    cf w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    cf a = buf[p];
    cf b = cmulc(buf[p + kSP4096], w1);
    cf c = cmulc(buf[p + 2 * kSP4096], w2);
    cf d = cmulc(buf[p + 3 * kSP4096], w3);
    fft4<+1>(a, b, c, d);
    buf[p] = a;
    buf[p + kSP4096] = b;
    buf[p + 2 * kSP4096] = c;
    buf[p + 3 * kSP4096] = d;
  }
  __syncthreads();
}

__device__ __forceinline__ int rev_big(int k) {
  return ((k & 3) << 12) | (((k >> 2) & 15) << 8) | (((k >> 6) & 15) << 4) | ((k >> 10) & 15);
}

// Spectrum of the packed real signal at bin k (0 <= k <= kM) from the
// digit-reversed half-size complex spectrum in buf:  X[k] = E + W^k O.
__device__ __forceinline__ cf unpack_bin(cf zk, cf zmk_conj, cf wk) {
  cf e = make_float2(0.5f * (zk.x + zmk_conj.x), 0.5f * (zk.y + zmk_conj.y));
  cf d = csub(zk, zmk_conj);                      // (Zk - conj Zmk)
  cf o = make_float2(0.5f * d.y, -0.5f * d.x);    // d / (2i)
  return cadd(e, cmul(wk, o));
}

struct EnvTerm {
  const float* src;   // clip base pointer
  long long base;     // absolute sample index of output sample 0 inside the clip
  const float* rir;   // interleaved (L, 2)
  int L;
};

// Fill buf with the packed circular source sequence u (see header comment):
//   u[m] = src[base + m]      0 <= m < sr
//   u[P - j] = src[base - j]  1 <= j <= L-1  (0 where base - j < 0)
__device__ void load_source(cf* buf, const EnvTerm& t, int sr) {
  for (int n = threadIdx.x; n < kM; n += kThreads) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int m = 2 * n + h;
      float x = 0.f;
      if (m < sr) {
        x = __ldg(t.src + t.base + m);
      } else if (m > kP - t.L) {
        long long idx = t.base - (kP - m);
        if (idx >= 0) x = __ldg(t.src + idx);
      }
      v[h] = x;
    }
    buf[padi(n)] = make_float2(v[0], v[1]);
  }
  __syncthreads();
}

__device__ void load_rir(cf* buf, const EnvTerm& t, int c) {
  for (int n = threadIdx.x; n < kM; n += kThreads) {
    int k0 = 2 * n, k1 = 2 * n + 1;
    float a = (k0 < t.L) ? __ldg(t.rir + 2 * (long long)k0 + c) : 0.f;
    float b = (k1 < t.L) ? __ldg(t.rir + 2 * (long long)k1 + c) : 0.f;
    buf[padi(n)] = make_float2(a, b);
  }
  __syncthreads();
}

// buf holds the digit-reversed spectrum of a packed real signal.  Write its
// real-signal spectrum X[0..kM] in natural order to xs (global scratch).
__device__ void store_spectrum(const cf* buf, const cf* __restrict__ tw, cf* xs) {
  for (int k = threadIdx.x; k <= kM / 2; k += kThreads) {
    if (k == 0) {
      cf z = buf[padi(0)];
      __stcg(&xs[0], make_float2(z.x + z.y, 0.f));
      __stcg(&xs[kM], make_float2(z.x - z.y, 0.f));
    } else {
      cf zk = buf[padi(rev_big(k))], zm = buf[padi(rev_big(kM - k))];
      cf wk = tw[k];
      cf xk = unpack_bin(zk, cconj(zm), wk);
      // W^{M-k} = -conj(W^k)
      cf xm = unpack_bin(zm, cconj(zk), make_float2(-wk.x, wk.y));
      __stcg(&xs[k], xk);
      if (k != kM / 2) __stcg(&xs[kM - k], xm);
    }
  }
  __syncthreads();
}

// buf holds the digit-reversed half-size spectrum of the packed RIR channel.
// Multiply with the source spectrum xs (natural order, global), optionally add
// / store the running sum yacc, and (when `finish`) repack the product in place
// as the half-size spectrum of the packed real result, scaled by 1/M.
__device__ void spectral_multiply(cf* buf, const cf* __restrict__ tw, const cf* xs, cf* yacc,
                                  bool add_acc, bool finish) {
  const float scale = 1.0f / (float)kM;  // inverse half-size FFT is unnormalised
  for (int k = threadIdx.x; k <= kM / 2; k += kThreads) {
    cf yk, ym, wk;
    if (k == 0) {
      cf z = buf[padi(0)];
      cf x0 = __ldcg(&xs[0]), xM = __ldcg(&xs[kM]);
      yk = make_float2((z.x + z.y) * x0.x, 0.f);   // Y[0]  (real)
      ym = make_float2((z.x - z.y) * xM.x, 0.f);   // Y[M]  (real)
      wk = make_float2(1.f, 0.f);
    } else {
      cf zk = buf[padi(rev_big(k))], zm = buf[padi(rev_big(kM - k))];
      wk = tw[k];
      cf ak = unpack_bin(zk, cconj(zm), wk);
      cf am = unpack_bin(zm, cconj(zk), make_float2(-wk.x, wk.y));
      yk = cmul(ak, __ldcg(&xs[k]));
      ym = cmul(am, __ldcg(&xs[kM - k]));
    }
    if (add_acc) {
      yk = cadd(yk, __ldcg(&yacc[k]));
      ym = cadd(ym, __ldcg(&yacc[kM - k]));
    }
    if (!finish) {
      __stcg(&yacc[k], yk);
      if (k != kM / 2) __stcg(&yacc[kM - k], ym);
      continue;
    }
    // repack: Zy[k] = E + i O, E = (Y[k] + conj Y[M-k])/2, O = (Y[k] - conj Y[M-k])/2 * conj(W^k)
    if (k == 0) {
      float e = 0.5f * (yk.x + ym.x), o = 0.5f * (yk.x - ym.x);
      buf[padi(0)] = make_float2(e * scale, o * scale);
    } else {
      cf ymc = cconj(ym);
      cf e = make_float2(0.5f * (yk.x + ymc.x), 0.5f * (yk.y + ymc.y));
      cf d = make_float2(0.5f * (yk.x - ymc.x), 0.5f * (yk.y - ymc.y));
      cf o = cmulc(d, wk);
      buf[padi(rev_big(k))] = make_float2((e.x - o.y) * scale, (e.y + o.x) * scale);
      if (k != kM / 2) {
        // Zy[M-k]: E' = conj(E), O' = (Y[M-k] - conj Y[k])/2 * conj(W^{M-k}) = -conj(d) * (-W^k) = conj(d) * W^k
        cf o2 = cmul(cconj(d), wk);
        buf[padi(rev_big(kM - k))] = make_float2((e.x - o2.y) * scale, (-e.y + o2.x) * scale);
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------- STFT part

struct SmemSamples {  // waveform held in the big FFT buffer (packed, padded)
  const float* f;
  __device__ __forceinline__ float operator()(int i) const { return f[2 * padi(i >> 1) + (i & 1)]; }
  // samples (i, i + 1), i even: one packed element
  __device__ __forceinline__ float2 pair(int i) const { return reinterpret_cast<const float2*>(f)[padi(i >> 1)]; }
};
struct GmemSamples {  // rows of an (N, 2, sr) tensor: 8-byte aligned when sr is even (checked by the host)
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return __ldg(p + i); }
  __device__ __forceinline__ float2 pair(int i) const { return __ldg(reinterpret_cast<const float2*>(p + i)); }
};

// Shared-memory tables behind the 512-float window (filled by fill_window):
//   st1[q * 16 + l] = w_256^(l q)   the twiddles of the first radix-16 stage of a frame (constant per lane)
//   stw[k]          = w_512^k       k = 0 .. 128, the real-spectrum unpacking factors
constexpr int kStftTabElems = 256 + 136;
constexpr size_t kStftSmem = kNfft * sizeof(float) + kStftTabElems * sizeof(cf);

__device__ __forceinline__ float sqrt_approx(float x) {
#ifdef AVL_HOST_EMUL
  return sqrtf(x);
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

// r16_fwd with the twiddles w^q read from a table (stride 16 between q's)
template <int SP>
__device__ __forceinline__ void r16_fwd_tab(cf* buf, int p0, const cf* tab) {
  cf x[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) x[m] = buf[p0 + m * SP];
  fft16<-1>(x);
#pragma unroll
  for (int q = 1; q < 16; ++q) x[o16(q)] = cmul(x[o16(q)], tab[16 * q]);
#pragma unroll
  for (int q = 0; q < 16; ++q) buf[p0 + q * SP] = x[o16(q)];
}

// STFT -> |.| -> 4x4 zero-padded block mean -> log1p for one channel.
// fb: per-warp frame buffers (2 * kFrameElems float2 per warp); win: 512-float window in smem followed by the tables above.
// out: spectrogram base pointer for this env, layout (65, TB, 2); writes channel c.
// Warp-local: a warp owns whole time blocks (4 frames, two at a time — one per half-warp), sums the magnitudes of its
// four frames in registers and pools the 65 frequency blocks itself, so the warps of a CTA run out of phase and hide
// one another's latencies; the only CTA-wide barrier is the one that releases the waveform at the end.
template <class Samples>
__device__ void stft_channel(const Samples& y, int sr, cf* fb, const float* win,
                             const cf* __restrict__ tw, float* out, int c) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, l = lane & 15;
  const int n_frames = 1 + sr / kHop;
  const int n_tb = (n_frames + 3) >> 2;
  const bool pair_ok = (sr & 1) == 0;  // rows of the waveform tensor stay 8-byte aligned
  const float2* win2 = reinterpret_cast<const float2*>(win);
  const cf* st1 = reinterpret_cast<const cf*>(win + kNfft) + l;
  const cf* stw = reinterpret_cast<const cf*>(win + kNfft) + 256;
  cf* z = fb + (warp * 2 + half) * kFrameElems;
  float* mag = reinterpret_cast<float*>(fb + warp * 2 * kFrameElems);  // 257 floats, after the frames are consumed
  (void)tw;
  for (int tb = warp; tb < n_tb; tb += kWarps) {
    float acc[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) acc[i] = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      if (4 * tb + 2 * pass >= n_frames) break;  // warp-uniform: both frames of the pass are padding
      const int f = 4 * tb + 2 * pass + half;
      const bool valid = f < n_frames;
      // windowed frame, packed as 256 complex: element j = (samples 2j, 2j+1); the window's support is j in [28, 228)
      const int start = kHop * f - kNfft / 2;  // even
      const bool interior = valid && pair_ok && start + (kNfft - kWin) / 2 >= 0 && start + (kNfft + kWin) / 2 <= sr;
      if (interior) {
        z[l] = make_float2(0.f, 0.f);
        z[l + 17 * 15] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 1; t < 15; ++t) {
          const int j = l + 16 * t;
          float2 v = make_float2(0.f, 0.f);
          if ((t > 1 && t < 14) || (t == 1 && l >= 12) || (t == 14 && l < 4)) {
            const float2 s2 = y.pair(start + 2 * j), w2 = win2[j];
            v = make_float2(w2.x * s2.x, w2.y * s2.y);
          }
          z[l + 17 * t] = v;
        }
      } else {
#pragma unroll 4
        for (int t = 0; t < 16; ++t) {
          const int j = l + 16 * t;
          const int i = 2 * j;
          float v0 = 0.f, v1 = 0.f;
          if (valid && i >= (kNfft - kWin) / 2 && i < (kNfft + kWin) / 2) {  // reflect padding at the clip's ends
            int i0 = start + i, i1 = start + i + 1;
            if (i0 < 0) i0 = -i0;
            if (i0 >= sr) i0 = 2 * (sr - 1) - i0;
            if (i1 < 0) i1 = -i1;
            if (i1 >= sr) i1 = 2 * (sr - 1) - i1;
            v0 = win[i] * y(i0);
            v1 = win[i + 1] * y(i1);
          }
          z[l + 17 * t] = make_float2(v0, v1);
        }
      }
      __syncwarp();
      // 256-point complex FFT = 16 x 16, one radix-16 butterfly per lane per stage
      r16_fwd_tab<kSP16>(z, l, st1);  // points l + 16 m -> padf = l + 17 m;  twiddles w_256^(l q)
      __syncwarp();
      r16_fwd<1>(z, 17 * l, false, make_float2(1.f, 0.f));  // points 16 l + m -> padf = 17 l + m
      __syncwarp();
      // real spectrum magnitudes: bins k and 256-k, k = 1..128 (8 per lane), plus 0 and 256
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int k = 1 + l + 16 * t;  // 1..128
        const int km = 256 - k;
        const cf zk = z[padf(((k & 15) << 4) | (k >> 4))];
        const cf zm = z[padf(((km & 15) << 4) | (km >> 4))];
        const cf wk = stw[k];  // exp(-2 pi i k / 512)
        const cf xk = unpack_bin(zk, cconj(zm), wk);
        const cf xm = unpack_bin(zm, cconj(zk), make_float2(-wk.x, wk.y));
        acc[2 * t] += sqrt_approx(fmaf(xk.x, xk.x, xk.y * xk.y));
        acc[2 * t + 1] += sqrt_approx(fmaf(xm.x, xm.x, xm.y * xm.y));
      }
      const cf z0 = z[0];
      acc[16] += fabsf(z0.x + z0.y);
      acc[17] += fabsf(z0.x - z0.y);
      __syncwarp();
    }
    // the two half-warps hold the sums of frames (0, 2) and (1, 3) of the block
#pragma unroll
    for (int i = 0; i < 18; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    if (half == 0) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int k = 1 + l + 16 * t;
        mag[k] = acc[2 * t];
        if (k != 128) mag[256 - k] = acc[2 * t + 1];
      }
      if (l == 0) {
        mag[0] = acc[16];
        mag[256] = acc[17];
      }
    }
    __syncwarp();
    // 4x4 block mean (zero padded in both directions) + log1p
    for (int kb = lane; kb < kFB; kb += 32) {
      float s = 0.f;
#pragma unroll
      for (int dk = 0; dk < 4; ++dk) {
        const int k = 4 * kb + dk;
        if (k < kBins) s += mag[k];
      }
      out[(kb * n_tb + tb) * 2 + c] = log1pf(s * 0.0625f);
    }
    __syncwarp();
  }
  __syncthreads();  // every warp is done with the waveform and the frame buffers
}

__device__ void fill_window(float* win, const cf* __restrict__ tw) {
  for (int i = threadIdx.x; i < kNfft; i += kThreads) {
    int n = i - (kNfft - kWin) / 2;
    float w = 0.f;
    if (n >= 0 && n < kWin) w = (float)(0.5 - 0.5 * cospi(2.0 * (double)n / (double)kWin));
    win[i] = w;
  }
  // tw[k] = exp(-2 pi i k / 32768):  w_256^m = tw[128 m],  w_512^k = tw[64 k]
  cf* tab = reinterpret_cast<cf*>(win + kNfft);
  for (int i = threadIdx.x; i < 256; i += kThreads) {
    const int m = (i >> 4) * (i & 15);  // <= 225; the table stops at half a turn: w_256^m = -w_256^(m - 128)
    const cf w = tw[128 * (m & 127)];
    tab[i] = m < 128 ? w : make_float2(-w.x, -w.y);
  }
  for (int i = threadIdx.x; i <= 128; i += kThreads) tab[256 + i] = tw[64 * i];
}

struct RenderArgs {
  int n_envs, sr;
  const float* sounds;
  const long long* clip_off;   // start of the env's clip inside `sounds`
  const int* index;            // _audio_index (seconds into the clip)
  const float* rirs;
  const long long* rir_off;    // in frames (2 floats per frame)
  const int* rir_len;
  const int* silent;
  // distractor (all null when absent)
  const long long* d_clip_off;
  const long long* d_rir_off;
  const int* d_rir_len;
  float* audiogoal;            // (N, 2, sr) or null
  float* spectrogram;          // (N, 65, TB, 2)
  const cf* tw;
  cf* scratch;                 // per CTA: 3 * (kM + 1) float2
  int* status;
  int split;                   // 1: one CTA per (env, ear) — rollout batches that would leave most SMs idle; the source
                               // spectrum is then computed by both CTAs of an env (3 big FFTs + 1 STFT each instead of
                               // 5 + 2 in one CTA): ~0.6x the latency for 1.2x the work
};

__global__ void __launch_bounds__(kThreads, 1) audio_render_kernel(RenderArgs a) {
  AVL_DYN_SMEM(smem_raw);
  cf* buf = reinterpret_cast<cf*>(smem_raw);
  cf* fb = buf + kBufElems;
  float* win = reinterpret_cast<float*>(fb + kWarps * 2 * kFrameElems);
  fill_window(win, a.tw);
  cf* xs0 = a.scratch + (size_t)blockIdx.x * 3 * (kM + 1);
  cf* xs1 = xs0 + (kM + 1);
  cf* yacc = xs1 + (kM + 1);
  const int sr = a.sr;
  const int n_tb = ((1 + sr / kHop) + 3) >> 2;
  const int lmax = kP - sr + 1;
  __syncthreads();

  const int items = a.split ? 2 * a.n_envs : a.n_envs;
  for (int w = blockIdx.x; w < items; w += gridDim.x) {
    const int e = a.split ? (w >> 1) : w;
    const int c_first = a.split ? (w & 1) : 0, c_last = a.split ? (w & 1) : 1;
    EnvTerm term[2];
    int nterms = 0;
    const bool silent = a.silent[e] != 0;
    if (!silent) {
      int L = a.rir_len[e];
      if (L > lmax) { L = lmax; if (threadIdx.x == 0) atomicExch(a.status, 1); }
      if (L > 0) {
        term[nterms].src = a.sounds + a.clip_off[e];
        term[nterms].base = (long long)a.index[e] * sr;
        term[nterms].rir = a.rirs + 2 * a.rir_off[e];
        term[nterms].L = L;
        ++nterms;
      }
      if (a.d_clip_off != nullptr) {
        int Ld = a.d_rir_len[e];
        if (Ld > lmax) { Ld = lmax; if (threadIdx.x == 0) atomicExch(a.status, 1); }
        if (Ld > 0) {
          term[nterms].src = a.sounds + a.d_clip_off[e];
          term[nterms].base = 0;
          term[nterms].rir = a.rirs + 2 * a.d_rir_off[e];
          term[nterms].L = Ld;
          ++nterms;
        }
      }
    }
    float* spec = a.spectrogram + (size_t)e * kFB * n_tb * 2;
    if (nterms == 0) {  // exact zeros: log1p(0) = 0 (belief_predictor.py:159 relies on it)
      if (c_first == 0) {  // (split mode: the CTA of the first ear clears both)
        for (int i = threadIdx.x; i < kFB * n_tb * 2; i += kThreads) spec[i] = 0.f;
        if (a.audiogoal) {
          float* ag = a.audiogoal + (size_t)e * 2 * sr;
          for (int i = threadIdx.x; i < 2 * sr; i += kThreads) ag[i] = 0.f;
        }
      }
      continue;
    }
    for (int t = 0; t < nterms; ++t) {
      load_source(buf, term[t], sr);
      fft_big_fwd(buf, a.tw);
      store_spectrum(buf, a.tw, t == 0 ? xs0 : xs1);
    }
    for (int c = c_first; c <= c_last; ++c) {
      for (int t = 0; t < nterms; ++t) {
        load_rir(buf, term[t], c);
        fft_big_fwd(buf, a.tw);
        spectral_multiply(buf, a.tw, t == 0 ? xs0 : xs1, yacc, t > 0, t == nterms - 1);
      }
      fft_big_inv(buf, a.tw);
      SmemSamples y{reinterpret_cast<const float*>(buf)};
      if (a.audiogoal) {
        float* ag = a.audiogoal + ((size_t)e * 2 + c) * sr;
        for (int i = threadIdx.x; i < sr; i += kThreads) ag[i] = y(i);
      }
      stft_channel(y, sr, fb, win, a.tw, spec, c);
    }
  }
}

// Stand-alone SpectrogramSensor.compute_spectrogram for a batch of waveforms (N, 2, sr).
__global__ void __launch_bounds__(kThreads, 1)
spectrogram_kernel(const float* audio, int n, int sr, float* spectrogram, const cf* tw) {
  AVL_DYN_SMEM(smem_raw);
  cf* fb = reinterpret_cast<cf*>(smem_raw);
  float* win = reinterpret_cast<float*>(fb + kWarps * 2 * kFrameElems);
  fill_window(win, tw);
  __syncthreads();
  const int n_tb = ((1 + sr / kHop) + 3) >> 2;
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    for (int c = 0; c < 2; ++c) {
      GmemSamples y{audio + ((size_t)e * 2 + c) * sr};
      stft_channel(y, sr, fb, win, tw, spectrogram + (size_t)e * kFB * n_tb * 2, c);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Spectral asset banks.  Of the transforms of one rendering only the inverse depends on the step: the spectrum of a
// scene's RIR (azimuth, receiver, source) and the spectrum of second `index` of a sound are properties of the assets.
// With 180 GB of HBM they can stay resident next to (or instead of) the time-domain banks: 16385 complex bins per real
// signal of the 32768-point circular convolution (128 KB; a binaural RIR 256 KB; a 28-node scene's 3136 RIRs 0.8 GB,
// built in ~1 ms of GPU time per 148 transforms).  A rendering is then, per (env, ear):
//     Y[k] = sum_terms A[k] X[k]  ->  one inverse FFT  ->  STFT / |.| / 4x4 mean / log1p
// i.e. 1 big transform instead of 2.5 and no shared work between the ears, so every (env, ear) is its own work item at
// every batch size.  The source row carries the whole history a RIR of the maximum length can reach
// (u[P - j] = src[base - j], 1 <= j <= P - sr): shorter RIRs multiply the extra history by zero taps.
struct SpectraArgs {
  int n, sr, kind;            // kind 0: RIRs (two items per row: the ears), 1: source seconds
  const float* bank;          // rirs (interleaved (L, 2)) / sounds
  const long long* off;       // rir_off (frames) / clip_off (samples)
  const int* len_or_index;    // rir_len / index (second of the clip)
  cf* out;                    // kind 0: (n, 2, kM + 1), kind 1: (n, kM + 1); natural order
  const cf* tw;
  int* status;
};

__global__ void __launch_bounds__(kThreads, 1) audio_spectra_kernel(SpectraArgs a) {
  AVL_DYN_SMEM(smem_raw);
  cf* buf = reinterpret_cast<cf*>(smem_raw);
  const int lmax = kP - a.sr + 1;
  const int items = a.kind == 0 ? 2 * a.n : a.n;
  for (int w = blockIdx.x; w < items; w += gridDim.x) {
    EnvTerm t;
    if (a.kind == 0) {
      const int r = w >> 1;
      int L = a.len_or_index[r];
      if (L > lmax) { L = lmax; if (threadIdx.x == 0) atomicExch(a.status, 1); }
      t.src = nullptr; t.base = 0; t.rir = a.bank + 2 * a.off[r]; t.L = L < 0 ? 0 : L;
      load_rir(buf, t, w & 1);
    } else {
      t.src = a.bank + a.off[w]; t.base = (long long)a.len_or_index[w] * a.sr; t.rir = nullptr; t.L = lmax;
      load_source(buf, t, a.sr);
    }
    fft_big_fwd(buf, a.tw);
    store_spectrum(buf, a.tw, a.out + (size_t)w * (kM + 1));
  }
}

struct SpectralArgs {
  int n_envs, sr;
  const cf* src_spec;          // rows of kM + 1 bins
  const long long* src_row0;   // row of second 0 of the env's clip; the rendered row is src_row0 + index
  const int* index;
  const cf* rir_spec;          // rows of (2, kM + 1) bins
  const long long* rir_row;    // < 0: empty RIR file
  const int* silent;
  const long long* d_src_row0; // distractor (null when absent): always second 0 of its clip (simulator.py:682-697)
  const long long* d_rir_row;
  float* audiogoal;            // (N, 2, sr) or null
  float* spectrogram;          // (N, 65, TB, 2)
  const cf* tw;
};

// buf <- half-size spectrum (digit-reversed, scaled by 1 / M) of the packed real signal with spectrum sum_t A_t X_t
__device__ void spectral_combine(cf* buf, const cf* __restrict__ tw, const cf* A0, const cf* X0, const cf* A1, const cf* X1) {
  const float scale = 1.0f / (float)kM;
  auto place = [&](int k, cf yk, cf ym) {
    if (k == 0) {  // Y[0], Y[M] are real
      const float e = 0.5f * (yk.x + ym.x), o = 0.5f * (yk.x - ym.x);
      buf[padi(0)] = make_float2(e * scale, o * scale);
      return;
    }
    const cf wk = tw[k];
    const cf ymc = cconj(ym);
    const cf e = make_float2(0.5f * (yk.x + ymc.x), 0.5f * (yk.y + ymc.y));
    const cf d = make_float2(0.5f * (yk.x - ymc.x), 0.5f * (yk.y - ymc.y));
    const cf o = cmulc(d, wk);
    buf[padi(rev_big(k))] = make_float2((e.x - o.y) * scale, (e.y + o.x) * scale);
    if (k != kM / 2) {
      const cf o2 = cmul(cconj(d), wk);
      buf[padi(rev_big(kM - k))] = make_float2((e.x - o2.y) * scale, (-e.y + o2.x) * scale);
    }
  };
  // the spectrum rows come from HBM / L2: the loads of four bins (and their mirrors) are issued before the first product
  // (profiles/r02_audio_spectral_ncu_full.txt: 3.6 long-scoreboard stalls per issued instruction with one bin in flight)
  constexpr int B = 4;
  int k = threadIdx.x;
  for (; k + (B - 1) * kThreads <= kM / 2; k += B * kThreads) {
    cf ak[B], xk[B], am[B], xm[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int kk = k + j * kThreads;
      ak[j] = __ldg(A0 + kk); xk[j] = __ldg(X0 + kk);
      am[j] = __ldg(A0 + kM - kk); xm[j] = __ldg(X0 + kM - kk);
    }
    cf yk[B], ym[B];
#pragma unroll
    for (int j = 0; j < B; ++j) { yk[j] = cmul(ak[j], xk[j]); ym[j] = cmul(am[j], xm[j]); }
    if (A1 != nullptr) {  // distractor term
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const int kk = k + j * kThreads;
        ak[j] = __ldg(A1 + kk); xk[j] = __ldg(X1 + kk);
        am[j] = __ldg(A1 + kM - kk); xm[j] = __ldg(X1 + kM - kk);
      }
#pragma unroll
      for (int j = 0; j < B; ++j) { yk[j] = cadd(yk[j], cmul(ak[j], xk[j])); ym[j] = cadd(ym[j], cmul(am[j], xm[j])); }
    }
#pragma unroll
    for (int j = 0; j < B; ++j) place(k + j * kThreads, yk[j], ym[j]);
  }
  for (; k <= kM / 2; k += kThreads) {
    cf yk = cmul(__ldg(A0 + k), __ldg(X0 + k));
    cf ym = cmul(__ldg(A0 + kM - k), __ldg(X0 + kM - k));
    if (A1 != nullptr) {
      yk = cadd(yk, cmul(__ldg(A1 + k), __ldg(X1 + k)));
      ym = cadd(ym, cmul(__ldg(A1 + kM - k), __ldg(X1 + kM - k)));
    }
    place(k, yk, ym);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) audio_render_spectral_kernel(SpectralArgs a) {
  AVL_DYN_SMEM(smem_raw);
  cf* buf = reinterpret_cast<cf*>(smem_raw);
  cf* fb = buf + kBufElems;
  float* win = reinterpret_cast<float*>(fb + kWarps * 2 * kFrameElems);
  fill_window(win, a.tw);
  const int sr = a.sr;
  const int n_tb = ((1 + sr / kHop) + 3) >> 2;
  __syncthreads();
  for (int w = blockIdx.x; w < 2 * a.n_envs; w += gridDim.x) {
    const int e = w >> 1, c = w & 1;
    const cf *A0 = nullptr, *X0 = nullptr, *A1 = nullptr, *X1 = nullptr;
    if (a.silent[e] == 0) {
      const long long r = a.rir_row[e];
      if (r >= 0) {
        A0 = a.rir_spec + (size_t)(2 * r + c) * (kM + 1);
        X0 = a.src_spec + (size_t)(a.src_row0[e] + (a.index ? a.index[e] : 0)) * (kM + 1);
      }
      if (a.d_src_row0 != nullptr) {
        const long long rd = a.d_rir_row[e];
        if (rd >= 0) {
          A1 = a.rir_spec + (size_t)(2 * rd + c) * (kM + 1);
          X1 = a.src_spec + (size_t)a.d_src_row0[e] * (kM + 1);
        }
      }
      if (A0 == nullptr) { A0 = A1; X0 = X1; A1 = nullptr; X1 = nullptr; }
    }
    float* spec = a.spectrogram + (size_t)e * kFB * n_tb * 2;
    if (A0 == nullptr) {  // exact zeros, as in audio_render_kernel
      for (int i = threadIdx.x; i < kFB * n_tb; i += kThreads) spec[2 * i + c] = 0.f;
      if (a.audiogoal) {
        float* ag = a.audiogoal + ((size_t)e * 2 + c) * sr;
        for (int i = threadIdx.x; i < sr; i += kThreads) ag[i] = 0.f;
      }
      continue;
    }
    spectral_combine(buf, a.tw, A0, X0, A1, X1);
    fft_big_inv(buf, a.tw);
    SmemSamples y{reinterpret_cast<const float*>(buf)};
    if (a.audiogoal) {
      float* ag = a.audiogoal + ((size_t)e * 2 + c) * sr;
      for (int i = threadIdx.x; i < sr; i += kThreads) ag[i] = y(i);
    }
    stft_channel(y, sr, fb, win, a.tw, spec, c);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Partitioned convolution for sampling rates / RIR lengths beyond one 32768-point circular convolution
// (sr + L - 1 > 32768: Replica's 44.1 kHz clips, nav.py:87-101 -> (65, 69, 2) spectrograms).  Uniformly partitioned
// overlap-save with blocks of B = 16384 samples on the SAME 16384-point complex FFT machinery:
//     y[k B + n] = sum_q  circ_P( h_q , u_{k-q} )[n],   n in [0, B),   P = 2 B
//     h_q[t] = rir[q B + t]  (t < B),        u_j[m] = src[base + j B + m]  (m < B),  u_j[P - i] = src[base + j B - i]  (i < B)
// Spectra X_j of the source windows and the accumulators Y_{c,k} live in a per-CTA global scratch (L2 resident); per
// (term, channel, partition) one forward FFT of the RIR partition is multiplied into every output block it reaches; one
// inverse FFT per (channel, block).  The waveform goes to global memory (the audiogoal output or a per-CTA scratch) and
// the STFT reads it back with coherent loads.
struct GmemSamplesCg {  // written earlier by this CTA: coherent loads (not the read-only path)
  const float* p;
  __device__ __forceinline__ float operator()(int i) const { return __ldcg(p + i); }
  __device__ __forceinline__ float2 pair(int i) const { return __ldcg(reinterpret_cast<const float2*>(p + i)); }
};

// packed window j of the source: see the formulas above.  lim = first sample index that must read as zero (base + sr)
__device__ void load_window(cf* buf, const float* src, long long w0, long long lim) {
  for (int n = threadIdx.x; n < kM; n += kThreads) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = 2 * n + h;
      long long idx = -1;
      if (m < kM) idx = w0 + m;
      else if (m > kP - kM) idx = w0 - (kP - m);
      v[h] = (idx >= 0 && idx < lim) ? __ldg(src + idx) : 0.f;
    }
    buf[padi(n)] = make_float2(v[0], v[1]);
  }
  __syncthreads();
}

// packed partition q of channel c of an interleaved (L, 2) RIR
__device__ void load_rir_part(cf* buf, const float* rir, int L, int c, int q) {
  const long long o = (long long)q * kM;
  for (int n = threadIdx.x; n < kM; n += kThreads) {
    const long long k0 = o + 2 * n, k1 = k0 + 1;
    const float a = (2 * n < kM && k0 < L) ? __ldg(rir + 2 * k0 + c) : 0.f;
    const float b = (2 * n + 1 < kM && k1 < L) ? __ldg(rir + 2 * k1 + c) : 0.f;
    buf[padi(n)] = make_float2(a, b);
  }
  __syncthreads();
}

// buf <- half-size spectrum (digit-reversed, scaled by 1 / M) of the packed real signal whose spectrum is Y[0..kM]
__device__ void load_repacked(cf* buf, const cf* __restrict__ tw, const cf* Y) {
  const float scale = 1.0f / (float)kM;
  for (int k = threadIdx.x; k <= kM / 2; k += kThreads) {
    if (k == 0) {
      const cf y0 = __ldcg(&Y[0]), yM = __ldcg(&Y[kM]);
      buf[padi(0)] = make_float2(0.5f * (y0.x + yM.x) * scale, 0.5f * (y0.x - yM.x) * scale);
      continue;
    }
    const cf yk = __ldcg(&Y[k]), ymc = cconj(__ldcg(&Y[kM - k]));
    const cf wk = tw[k];
    const cf e = make_float2(0.5f * (yk.x + ymc.x), 0.5f * (yk.y + ymc.y));
    const cf d = make_float2(0.5f * (yk.x - ymc.x), 0.5f * (yk.y - ymc.y));
    const cf o = cmulc(d, wk);
    buf[padi(rev_big(k))] = make_float2((e.x - o.y) * scale, (e.y + o.x) * scale);
    if (k != kM / 2) {
      const cf o2 = cmul(cconj(d), wk);
      buf[padi(rev_big(kM - k))] = make_float2((e.x - o2.y) * scale, (-e.y + o2.x) * scale);
    }
  }
  __syncthreads();
}

struct PartArgs {
  int K;         // output blocks: ceil(sr / B)
  int Qmax;      // RIR partitions the scratch is sized for
  cf* scratch;   // per CTA: (K + Qmax - 1) source spectra, then 2 K accumulators, (kM + 1) complex each
  float* wave;   // per CTA 2 * sr floats (used when no audiogoal output is requested)
};

__global__ void __launch_bounds__(kThreads, 1) audio_render_part_kernel(RenderArgs a, PartArgs pa) {
  AVL_DYN_SMEM(smem_raw);
  cf* buf = reinterpret_cast<cf*>(smem_raw);
  cf* fb = buf + kBufElems;
  float* win = reinterpret_cast<float*>(fb + kWarps * 2 * kFrameElems);
  fill_window(win, a.tw);
  const int sr = a.sr, K = pa.K, Qmax = pa.Qmax;
  const size_t spec = (size_t)(kM + 1);
  cf* Xs = pa.scratch + (size_t)blockIdx.x * (size_t)(K + Qmax - 1 + 2 * K) * spec;
  cf* Ys = Xs + (size_t)(K + Qmax - 1) * spec;
  const int n_tb = ((1 + sr / kHop) + 3) >> 2;
  __syncthreads();
  for (int e = blockIdx.x; e < a.n_envs; e += gridDim.x) {
    EnvTerm term[2];
    int nterms = 0;
    if (a.silent[e] == 0) {
      int L = a.rir_len[e];
      if (L > Qmax * kM) { L = Qmax * kM; if (threadIdx.x == 0) atomicExch(a.status, 1); }
      if (L > 0) {
        term[nterms].src = a.sounds + a.clip_off[e];
        term[nterms].base = (long long)a.index[e] * sr;
        term[nterms].rir = a.rirs + 2 * a.rir_off[e];
        term[nterms].L = L;
        ++nterms;
      }
      if (a.d_clip_off != nullptr) {
        int Ld = a.d_rir_len[e];
        if (Ld > Qmax * kM) { Ld = Qmax * kM; if (threadIdx.x == 0) atomicExch(a.status, 1); }
        if (Ld > 0) {
          term[nterms].src = a.sounds + a.d_clip_off[e];
          term[nterms].base = 0;
          term[nterms].rir = a.rirs + 2 * a.d_rir_off[e];
          term[nterms].L = Ld;
          ++nterms;
        }
      }
    }
    float* spec_out = a.spectrogram + (size_t)e * kFB * n_tb * 2;
    float* wave = a.audiogoal ? a.audiogoal + (size_t)e * 2 * sr : pa.wave + (size_t)blockIdx.x * 2 * sr;
    if (nterms == 0) {
      for (int i = threadIdx.x; i < kFB * n_tb * 2; i += kThreads) spec_out[i] = 0.f;
      if (a.audiogoal)
        for (int i = threadIdx.x; i < 2 * sr; i += kThreads) wave[i] = 0.f;
      continue;
    }
    for (int t = 0; t < nterms; ++t) {
      const int Q = (term[t].L + kM - 1) / kM;
      for (int j = -(Q - 1); j < K; ++j) {
        load_window(buf, term[t].src, term[t].base + (long long)j * kM, term[t].base + sr);
        fft_big_fwd(buf, a.tw);
        store_spectrum(buf, a.tw, Xs + (size_t)(j + Qmax - 1) * spec);
      }
      for (int c = 0; c < 2; ++c)
        for (int q = 0; q < Q; ++q) {
          load_rir_part(buf, term[t].rir, term[t].L, c, q);
          fft_big_fwd(buf, a.tw);
          for (int k = 0; k < K; ++k)  // (t, q) = (0, 0) is the first product every accumulator receives
            spectral_multiply(buf, a.tw, Xs + (size_t)(k - q + Qmax - 1) * spec, Ys + (size_t)(c * K + k) * spec,
                              t > 0 || q > 0, false);
        }
    }
    for (int c = 0; c < 2; ++c) {
      for (int k = 0; k < K; ++k) {
        load_repacked(buf, a.tw, Ys + (size_t)(c * K + k) * spec);
        fft_big_inv(buf, a.tw);
        SmemSamples y{reinterpret_cast<const float*>(buf)};
        const int n_out = min(kM, sr - k * kM);
        float* wc = wave + (size_t)c * sr + (size_t)k * kM;
        for (int i = threadIdx.x; i < n_out; i += kThreads) __stcg(wc + i, y(i));
        __syncthreads();
      }
      GmemSamplesCg yc{wave + (size_t)c * sr};
      stft_channel(yc, sr, fb, win, a.tw, spec_out, c);
    }
  }
}

__global__ void twiddle_init_kernel(cf* tw) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= kM) {
    double s, c;
    sincospi((double)k / (double)kM, &s, &c);
    tw[k] = make_float2((float)c, (float)(-s));
  }
}

#ifndef AVL_HOST_EMUL
struct AudioCtx {
  int sr;
  int grid;
  cf* tw;
  cf* scratch;
  int* status;
  // partitioned path (sr + L - 1 may exceed one 32768-point convolution): allocated when sr > 16769
  int part_K, part_Qmax;
  cf* part_scratch;
  float* part_wave;
};

constexpr int kPartQmax = 3;  // RIRs up to 3 * 16384 = 49152 samples (1.1 s at 44.1 kHz)

constexpr size_t kRenderSmem = (size_t)(kBufElems + kWarps * 2 * kFrameElems) * sizeof(cf) + kStftSmem;
constexpr size_t kSpectraSmem = (size_t)kBufElems * sizeof(cf);
constexpr size_t kSpecSmem = (size_t)(kWarps * 2 * kFrameElems) * sizeof(cf) + kStftSmem;

#endif  // AVL_HOST_EMUL

}  // namespace

#ifndef AVL_HOST_EMUL
// Creates the audio context (twiddle table + per-CTA spectrum scratch).  The
// only allocation the audio path ever makes.  sr >= 512; above 16769 Hz the partitioned-convolution path is set up.
AVL_API int avl_audio_create(int sr, void** handle) {
  if (!handle) return AVL_ERR_ARG;
  if (sr < kNfft) return AVL_ERR_UNSUPPORTED;
  AudioCtx* ctx = new AudioCtx();
  ctx->sr = sr;
  ctx->grid = avl_num_sms();
  AVL_CUDA_CHECK(cudaMalloc(&ctx->tw, sizeof(cf) * (kM + 1)));
  AVL_CUDA_CHECK(cudaMalloc(&ctx->scratch, sizeof(cf) * 3 * (kM + 1) * (size_t)ctx->grid));
  AVL_CUDA_CHECK(cudaMalloc(&ctx->status, sizeof(int)));
  AVL_CUDA_CHECK(cudaMemset(ctx->status, 0, sizeof(int)));
  twiddle_init_kernel<<<avl_div_up(kM + 1, 256), 256>>>(ctx->tw);
  AVL_LAUNCH_CHECK();
  ctx->part_K = 0; ctx->part_Qmax = 0; ctx->part_scratch = nullptr; ctx->part_wave = nullptr;
  if (sr > kP / 2 + 385) {
    ctx->part_K = (sr + kM - 1) / kM;
    ctx->part_Qmax = kPartQmax;
    const size_t spectra = (size_t)(ctx->part_K + ctx->part_Qmax - 1 + 2 * ctx->part_K);
    AVL_CUDA_CHECK(cudaMalloc(&ctx->part_scratch, sizeof(cf) * spectra * (kM + 1) * (size_t)ctx->grid));
    AVL_CUDA_CHECK(cudaMalloc(&ctx->part_wave, sizeof(float) * 2 * (size_t)sr * (size_t)ctx->grid));
    AVL_CUDA_CHECK(cudaFuncSetAttribute(audio_render_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRenderSmem));
  }
  AVL_CUDA_CHECK(cudaFuncSetAttribute(audio_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRenderSmem));
  AVL_CUDA_CHECK(cudaFuncSetAttribute(spectrogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpecSmem));
  AVL_CUDA_CHECK(cudaFuncSetAttribute(audio_spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpectraSmem));
  AVL_CUDA_CHECK(cudaFuncSetAttribute(audio_render_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRenderSmem));
  AVL_CUDA_CHECK(cudaDeviceSynchronize());
  *handle = ctx;
  return AVL_OK;
}

AVL_API int avl_audio_destroy(void* handle) {
  AudioCtx* ctx = static_cast<AudioCtx*>(handle);
  if (!ctx) return AVL_ERR_ARG;
  cudaFree(ctx->tw);
  cudaFree(ctx->scratch);
  cudaFree(ctx->status);
  if (ctx->part_scratch) cudaFree(ctx->part_scratch);
  if (ctx->part_wave) cudaFree(ctx->part_wave);
  delete ctx;
  return AVL_OK;
}

// Sticky device-side status of earlier render calls (synchronises the device):
// 0 ok, 1 = some RIR was longer than 32768 - sr + 1 samples (49152 on the partitioned path) and was truncated.
AVL_API int avl_audio_status(void* handle, int* status_out) {
  AudioCtx* ctx = static_cast<AudioCtx*>(handle);
  if (!ctx || !status_out) return AVL_ERR_ARG;
  AVL_CUDA_CHECK(cudaMemcpy(status_out, ctx->status, sizeof(int), cudaMemcpyDeviceToHost));
  return AVL_OK;
}

static int g_audio_split = 1;
// 1 (default): batches of at most half the SM count render each ear in its own CTA; 0: always one CTA per env.
AVL_API int avl_set_audio_channel_split(int on) {
  int old = g_audio_split;
  g_audio_split = on ? 1 : 0;
  return old;
}

// Fused rows A + B.  All pointers are device pointers.  See include/avlen_b200.h.
AVL_API int avl_audio_render_spectrogram(void* handle, int n_envs, const float* sounds, const long long* clip_off,
                                         const int* index, const float* rirs, const long long* rir_off,
                                         const int* rir_len, const int* silent, const long long* d_clip_off,
                                         const long long* d_rir_off, const int* d_rir_len, float* audiogoal_out,
                                         float* spectrogram_out, void* stream) {
  AudioCtx* ctx = static_cast<AudioCtx*>(handle);
  if (!ctx || n_envs < 0) return AVL_ERR_ARG;
  if (n_envs == 0) return AVL_OK;
  if (!sounds || !clip_off || !index || !rirs || !rir_off || !rir_len || !silent || !spectrogram_out) return AVL_ERR_ARG;
  if ((d_clip_off != nullptr) != (d_rir_off != nullptr) || (d_clip_off != nullptr) != (d_rir_len != nullptr)) return AVL_ERR_ARG;
  RenderArgs a;
  a.n_envs = n_envs; a.sr = ctx->sr; a.sounds = sounds; a.clip_off = clip_off; a.index = index;
  a.rirs = rirs; a.rir_off = rir_off; a.rir_len = rir_len; a.silent = silent;
  a.d_clip_off = d_clip_off; a.d_rir_off = d_rir_off; a.d_rir_len = d_rir_len;
  a.audiogoal = audiogoal_out; a.spectrogram = spectrogram_out; a.tw = ctx->tw; a.scratch = ctx->scratch;
  a.status = ctx->status;
  if (ctx->part_K > 0) {  // sr + L - 1 may exceed one 32768-point convolution: partitioned overlap-save
    a.split = 0;
    PartArgs pa;
    pa.K = ctx->part_K; pa.Qmax = ctx->part_Qmax; pa.scratch = ctx->part_scratch; pa.wave = ctx->part_wave;
    const int g = n_envs < ctx->grid ? n_envs : ctx->grid;
    audio_render_part_kernel<<<g, kThreads, kRenderSmem, (cudaStream_t)stream>>>(a, pa);
    AVL_LAUNCH_CHECK();
    return AVL_OK;
  }
  a.split = (2 * n_envs <= ctx->grid && g_audio_split) ? 1 : 0;
  const int items = a.split ? 2 * n_envs : n_envs;
  int grid = items < ctx->grid ? items : ctx->grid;
  audio_render_kernel<<<grid, kThreads, kRenderSmem, (cudaStream_t)stream>>>(a);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// Complex bins per spectrum row of the spectral banks (16385: bins 0 .. 16384 of the 32768-point transform).
AVL_API int avl_audio_spectrum_bins(void) { return kM + 1; }

static int audio_spectra(AudioCtx* ctx, int kind, int n, const float* bank, const long long* off, const int* len_or_index,
                         float* out, void* stream) {
  if (!ctx || n < 0) return AVL_ERR_ARG;
  if (ctx->part_K > 0) return AVL_ERR_UNSUPPORTED;  // one 32768-point convolution only (sr <= 16769)
  if (n == 0) return AVL_OK;
  if (!bank || !off || !len_or_index || !out) return AVL_ERR_ARG;
  SpectraArgs a;
  a.n = n; a.sr = ctx->sr; a.kind = kind; a.bank = bank; a.off = off; a.len_or_index = len_or_index;
  a.out = reinterpret_cast<cf*>(out); a.tw = ctx->tw; a.status = ctx->status;
  const int items = kind == 0 ? 2 * n : n;
  // 140 KB of shared memory per CTA: one CTA per SM, each walking over the items
  audio_spectra_kernel<<<items < ctx->grid ? items : ctx->grid, kThreads, kSpectraSmem, (cudaStream_t)stream>>>(a);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// Spectral RIR bank: row i of spectra_out (n, 2, bins, 2) = spectra of the two ears of RIR i (rir_len 0: zeros; longer
// than 32768 - sr + 1: truncated, avl_audio_status reports it).  Device pointers.
AVL_API int avl_audio_rir_spectra(void* handle, int n, const float* rirs, const long long* rir_off, const int* rir_len,
                                  float* spectra_out, void* stream) {
  return audio_spectra(static_cast<AudioCtx*>(handle), 0, n, rirs, rir_off, rir_len, spectra_out, stream);
}

// Spectral sound bank: row i of spectra_out (n, bins, 2) = spectrum of second index[i] of the clip at clip_off[i] together
// with the 32768 - sr samples before it (zeros before the clip's start).
AVL_API int avl_audio_source_spectra(void* handle, int n, const float* sounds, const long long* clip_off, const int* index,
                                     float* spectra_out, void* stream) {
  return audio_spectra(static_cast<AudioCtx*>(handle), 1, n, sounds, clip_off, index, spectra_out, stream);
}

// Rows A + B from the spectral banks: same result as avl_audio_render_spectrogram on the time-domain assets the rows
// were made from.  src_row0[i] + index[i] (index may be NULL: 0) is env i's source row; rir_row[i] < 0 = empty RIR file.
AVL_API int avl_audio_render_spectral(void* handle, int n_envs, const float* src_spectra, const long long* src_row0,
                                      const int* index, const float* rir_spectra, const long long* rir_row,
                                      const int* silent, const long long* d_src_row0, const long long* d_rir_row,
                                      float* audiogoal_out, float* spectrogram_out, void* stream) {
  AudioCtx* ctx = static_cast<AudioCtx*>(handle);
  if (!ctx || n_envs < 0) return AVL_ERR_ARG;
  if (ctx->part_K > 0) return AVL_ERR_UNSUPPORTED;
  if (n_envs == 0) return AVL_OK;
  if (!src_spectra || !src_row0 || !rir_spectra || !rir_row || !silent || !spectrogram_out) return AVL_ERR_ARG;
  if ((d_src_row0 != nullptr) != (d_rir_row != nullptr)) return AVL_ERR_ARG;
  SpectralArgs a;
  a.n_envs = n_envs; a.sr = ctx->sr; a.src_spec = reinterpret_cast<const cf*>(src_spectra); a.src_row0 = src_row0;
  a.index = index; a.rir_spec = reinterpret_cast<const cf*>(rir_spectra); a.rir_row = rir_row; a.silent = silent;
  a.d_src_row0 = d_src_row0; a.d_rir_row = d_rir_row; a.audiogoal = audiogoal_out; a.spectrogram = spectrogram_out;
  a.tw = ctx->tw;
  const int items = 2 * n_envs;
  audio_render_spectral_kernel<<<items < ctx->grid ? items : ctx->grid, kThreads, kRenderSmem, (cudaStream_t)stream>>>(a);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}

// Row B alone: (N, 2, sr) waveforms -> (N, 65, ceil((1 + sr/160)/4), 2).
AVL_API int avl_audio_spectrogram(void* handle, int n, const float* audio, float* spectrogram_out, void* stream) {
  AudioCtx* ctx = static_cast<AudioCtx*>(handle);
  if (!ctx || n < 0) return AVL_ERR_ARG;
  if (n == 0) return AVL_OK;
  if (!audio || !spectrogram_out) return AVL_ERR_ARG;
  int grid = n < 2 * ctx->grid ? n : 2 * ctx->grid;
  spectrogram_kernel<<<grid, kThreads, kSpecSmem, (cudaStream_t)stream>>>(audio, n, ctx->sr, spectrogram_out, ctx->tw);
  AVL_LAUNCH_CHECK();
  return AVL_OK;
}
#endif  // AVL_HOST_EMUL
