// Implicit-GEMM Conv1d family for the EBEN training step (fp32 SIMT path).
//
// One tile skeleton, four gather modes, all on (B, C, T) fp32 time-contiguous
// tensors exactly as the reference lays them out ("fp32 contiguous
// (B,C,T)"):
//   FWD     y[b,co,t]  = epi( sum_{ci,k} W[co,ci,k] * x[b,ci,map(t*s + k*d - pad)] )
//           (reference: every nn.Conv1d on the path, eben_generator.py:112-166,
//            eben_discriminator.py:66-157, melgan_discriminator.py:89-156; the
//            reflect halo of padding_mode="reflect"/nn.ReflectionPad1d is folded
//            into map() and never materialised)
//   DGRAD   dx[b,ci,u] = epi( sum_{co,k} W[co,ci,k] * dy[b,co,(u+pad-k*d)/s] ),
//           phase-decomposed over (u+pad) mod s so no zero-MACs are issued, with
//           the reflect-halo images of u folded in.  Also IS the forward of
//           nn.ConvTranspose1d (eben_generator.py:241-249).
//   WGRAD   dW[co,ci,k] += sum_{b,t} dy[b,co,t] * x[b,ci,map(t*s + k*d - pad)]
//           (split over the (b,t) reduction, fp32 atomics into the flat bucket)
//   SCATTER dx[b,ci,map(t*s+k*d-pad)] += sum_co W[co,ci,k] * dy[b,co,t]
//           (col2im form of dgrad; used where stride >> 1 and Cin is tiny: the
//            STFT-as-conv backward)
//
// The body is written as __host__ __device__ phase functions with an explicit
// thread id so that tests/emu can run the *same code* on the CPU (threads
// serialised between barriers) - there is no GPU in the build container.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define VBX_HD __host__ __device__ __forceinline__
#define VBX_UNROLL _Pragma("unroll")
#else
#define VBX_UNROLL
#define VBX_HD inline
struct float4 { float x, y, z, w; };
#endif

namespace vbx {

enum Mode { FWD = 0, DGRAD = 1, WGRAD = 2, SCATTER = 3 };

struct GemmP {
  int B, Cin, Cout, Tin, Tout, K, stride, dil, pad, refl, groups;
  int Cin_g, Cout_g;
  int mtiles;             // M tiles per group (blockIdx.y = g*mtiles + mt)
  int split;              // WGRAD: reduction elements handled by one blockIdx.z
  const float* W;         // FWD: W[co][ci][k]; DGRAD: Wt[g][ci][co][k]; SCATTER: Wk[g][(ci,k)][co]
  const float* X;         // FWD/WGRAD: x (B,Cin,Tin);  DGRAD/SCATTER: dy (B,Cout,Tout)
  const float* DY;        // WGRAD: dy (B,Cout,Tout)
  float* Y;               // FWD: y; DGRAD/SCATTER: dx; WGRAD: dW
  const float* bias;      // per output channel, nullable
  const float* res;       // residual, same shape as Y, nullable
  unsigned char* mask;    // optional: 1 where pre-activation > 0
  float slope;            // LeakyReLU slope on the output (1 = identity)
  float beta;             // Y = beta*Y_old + result (0 = overwrite)
  const float* gate;      // optional (vbx_epilogue): activation y the output is a gradient of; v *= y > 0 ? 1 : gate_slope
  const float* fm_other;  // optional, with gate: v += fm_coef[0]*sign(y - fm_other) - fm_coef[1]*sign(y) before the gate
  const float* fm_coef;
  float gate_slope;
  float* gate_dbias;      // optional, with gate (host side only: conv_plan.h GateArgs runs the reduction after the kernel)
};

// the gate stage of vbx_epilogue on one value (exact products: sign() is -1/0/1, so the sum below rounds exactly like
// the unfused  grad_fm = c1*sd - c2*sg;  g = dgrad + grad_fm;  g * lrelu'(y)  it replaces)
VBX_HD float gate_apply(float v, float y, bool fm, float other, float c1, float c2, float gslope) {
  if (fm) {
    const float d = y - other;
    const float sd = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    const float sg = y > 0.f ? 1.f : (y < 0.f ? -1.f : 0.f);
#ifdef __CUDA_ARCH__
    v = __fadd_rn(v, __fsub_rn(__fmul_rn(c1, sd), __fmul_rn(c2, sg)));
#else
    { volatile float f = c1 * sd - c2 * sg; v = v + f; }
#endif
  }
  return y > 0.f ? v : v * gslope;
}

struct Blk { int x, y, z; };

// position in the padded domain -> index into x, or -1 when it falls in the zero halo
VBX_HD int map_pos(int p, int Tin, int refl) {
  if (p < 0) {
    if (p < -refl) return -1;
    return -p;
  }
  if (p >= Tin) {
    if (p >= Tin + refl) return -1;
    return 2 * (Tin - 1) - p;
  }
  return p;
}

VBX_HD int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
VBX_HD int mod_pos(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// DGRAD bookkeeping shared by every thread of a block (uniform) ------------------
struct DgradPhase {
  int r;        // first u >= 0 of this phase
  int Up;       // number of u of this phase per batch item
};
VBX_HD DgradPhase dgrad_phase(const GemmP& P, int ph) {
  DgradPhase d;
  d.r = mod_pos(ph - P.pad, P.stride);
  d.Up = d.r < P.Tin ? (P.Tin - d.r + P.stride - 1) / P.stride : 0;
  return d;
}
struct DgradImg {
  int k0, kstep, ntaps, tstep, kd0_hi, phase;
};
// taps k = k0 + j*kstep (j < ntaps) are the ones with (k*dil) % stride == phase; for those
// (p + pad - k*dil)/stride = (p + pad)/stride - kd0_hi - j*tstep.
VBX_HD DgradImg dgrad_taps(const GemmP& P, int phase) {
  DgradImg I;
  I.phase = phase;
  int g = gcd_i(P.dil, P.stride);
  I.kstep = P.stride / g;
  I.tstep = P.dil / g;
  I.k0 = -1;
  for (int k = 0; k < I.kstep && k < P.K; ++k)
    if ((k * P.dil) % P.stride == I.phase) { I.k0 = k; break; }
  I.ntaps = I.k0 < 0 ? 0 : (P.K - 1 - I.k0) / I.kstep + 1;
  I.kd0_hi = I.k0 < 0 ? 0 : (I.k0 * P.dil) / P.stride;
  return I;
}
// image 0: p = u ; image 1: p = -u (left mirror) ; image 2: p = 2(Tin-1)-u (right mirror)
VBX_HD DgradImg dgrad_img(const GemmP& P, int r, int img) {
  int p0 = img == 0 ? r : (img == 1 ? -r : 2 * (P.Tin - 1) - r);
  return dgrad_taps(P, mod_pos(p0 + P.pad, P.stride));
}
// does a tile of `tile` consecutive columns starting at n_lo of phase `ph` touch mirror image `img`?
VBX_HD bool dgrad_tile_needs_img(const GemmP& P, int ph, int n_lo, int tile, int img) {
  if (img == 0) return true;
  if (P.refl == 0) return false;
  DgradPhase d = dgrad_phase(P, ph);
  int N = P.B * d.Up;
  int n_hi = n_lo + tile - 1;
  if (n_lo >= N) return false;
  if (n_hi >= N) n_hi = N - 1;
  if (n_lo / d.Up != n_hi / d.Up) return true;
  int u_lo = d.r + (n_lo % d.Up) * P.stride, u_hi = d.r + (n_hi % d.Up) * P.stride;
  if (img == 1) return u_lo <= P.refl && u_hi >= 1;
  return u_hi >= P.Tin - 1 - P.refl && u_lo <= P.Tin - 2;
}

template <int TM_, int TN_, int RM_, int RN_>
struct Cfg {
  static const int TM = TM_, TN = TN_, RM = RM_, RN = RN_;
  static const int KC = 16, NT = 256;
  static const int TY = TM / RM, TX = TN / RN;
  static const int LDA = TM + 4, LDB = TN + 4;
  static const int RMV = RM < 4 ? RM : 4, RMG = RM / RMV;
  static const int RNV = 4, RNG = RN / RNV;
  static const int AROWS = NT / KC;                 // rows of A covered per pass (A fast along kk)
  static const int BCOLS_K = TN / (NT / KC);        // columns per thread when B is fast along kk
  static const int BROWS_N = NT >= TN ? NT / TN : 1; // kk rows covered per pass when B is fast along n
  static_assert(TY * TX == NT, "thread layout");
  static_assert(TN % 4 == 0 && RN % 4 == 0, "RN");
  static_assert(NT % TN == 0 || TN % NT == 0, "TN");
};

// MODE: gather mode.  BK: B tile is loaded "fast along kk" (lanes walk the reduction
// index) - always for WGRAD, and for FWD when the stride is so large that walking
// t is uncoalesced (STFT-as-conv).
template <class C, int MODE, bool BK>
struct Tile {
  static const int TM = C::TM, TN = C::TN, RM = C::RM, RN = C::RN, KC = C::KC, NT = C::NT;
  static const int NBCOL = BK ? C::BCOLS_K : 1;

  struct TS {                      // per-thread state
    float acc[RM][RN];
    int g, m_base, n_base;
    int M, N, Kred;
    // A loader
    int a_kkl, a_m0;
    long long a_base;              // element offset of row 0 of this tile in W
    int a_co, a_j;                 // DGRAD: decoded (co, j) of this thread's kk in the current chunk
    // B loader (fast along n): one column per thread
    int b_nl, b_r0;
    bool b_valid;
    long long b_base;              // element offset of (b, group base channel, 0)
    int b_t;                       // FWD: t*s - pad ; DGRAD: T0
    int b_c, b_k;                  // running (ci,k) / (co,j)
    // B loader (fast along kk): NBCOL columns per thread
    int bk_kkl, bk_n0;
    long long bk_off[NBCOL];       // FWD-BK: per-column x base ; WGRAD: ci*Tin
    int bk_t[NBCOL];               // FWD-BK: t*s - pad ; WGRAD: k*d - pad  (INT_MIN/2 => invalid col)
    int bk_c, bk_k;                // FWD-BK running (ci,k)
    // DGRAD uniform info
    DgradPhase ph;
    DgradImg im;
    int img;
    int red_lo, red_hi;            // WGRAD reduction range
  };

  static const int INVALID = -(1 << 30);

  // ---------------------------------------------------------------- prologue
  static VBX_HD void prologue(const GemmP& P, Blk blk, int tid, TS& s) {
    VBX_UNROLL
    for (int i = 0; i < RM; ++i) {
      VBX_UNROLL
      for (int j = 0; j < RN; ++j) s.acc[i][j] = 0.f;
    }
    s.g = blk.y / P.mtiles;
    s.m_base = (blk.y % P.mtiles) * TM;
    s.n_base = blk.x * TN;
    s.a_kkl = tid % KC;
    s.a_m0 = tid / KC;
    s.b_nl = tid % TN;
    s.b_r0 = tid / TN;
    s.bk_kkl = tid % KC;
    s.bk_n0 = tid / KC;
    s.img = 0;
    if (MODE == FWD) {
      s.M = P.Cout_g; s.N = P.B * P.Tout; s.Kred = P.Cin_g * P.K;
      s.a_base = (long long)(s.g * P.Cout_g + s.m_base) * s.Kred;
      if (!BK) {
        int n = s.n_base + s.b_nl;
        s.b_valid = n < s.N;
        int b = s.b_valid ? n / P.Tout : 0, t = s.b_valid ? n % P.Tout : 0;
        s.b_base = ((long long)b * P.Cin + s.g * P.Cin_g) * P.Tin;
        s.b_t = t * P.stride - P.pad;
        s.b_c = s.b_r0 / P.K; s.b_k = s.b_r0 % P.K;
      } else {
        VBX_UNROLL
        VBX_UNROLL
      for (int c = 0; c < NBCOL; ++c) {
          int n = s.n_base + s.bk_n0 + c * (NT / KC);
          if (n < s.N) {
            int b = n / P.Tout, t = n % P.Tout;
            s.bk_off[c] = ((long long)b * P.Cin + s.g * P.Cin_g) * P.Tin;
            s.bk_t[c] = t * P.stride - P.pad;
          } else { s.bk_off[c] = 0; s.bk_t[c] = INVALID; }
        }
        s.bk_c = s.bk_kkl / P.K; s.bk_k = s.bk_kkl % P.K;
      }
    } else if (MODE == DGRAD) {
      s.M = P.Cin_g;
      s.ph = dgrad_phase(P, blk.z);
      s.N = P.B * s.ph.Up;
      s.a_base = (long long)(s.g * P.Cin_g + s.m_base) * P.Cout_g * P.K;
      s.Kred = 0;
    } else if (MODE == WGRAD) {
      s.M = P.Cout_g; s.N = P.Cin_g * P.K;
      int tot = P.B * P.Tout;
      s.red_lo = blk.z * P.split;
      s.red_hi = s.red_lo + P.split < tot ? s.red_lo + P.split : tot;
      s.Kred = s.red_hi > s.red_lo ? s.red_hi - s.red_lo : 0;
      VBX_UNROLL
      for (int c = 0; c < NBCOL; ++c) {
        int n = s.n_base + s.bk_n0 + c * (NT / KC);
        if (n < s.N) {
          int ci = n / P.K, k = n % P.K;
          s.bk_off[c] = (long long)(s.g * P.Cin_g + ci) * P.Tin;
          s.bk_t[c] = k * P.dil - P.pad;
        } else { s.bk_off[c] = 0; s.bk_t[c] = INVALID; }
      }
    } else {  // SCATTER
      s.M = P.Cin_g * P.K; s.N = P.B * P.Tout; s.Kred = P.Cout_g;
      s.a_base = ((long long)s.g * s.M + s.m_base) * P.Cout_g;
      int n = s.n_base + s.b_nl;
      s.b_valid = n < s.N;
      int b = s.b_valid ? n / P.Tout : 0, t = s.b_valid ? n % P.Tout : 0;
      s.b_base = ((long long)b * P.Cout + s.g * P.Cout_g) * P.Tout + t;
      s.b_t = 0; s.b_c = s.b_r0; s.b_k = 0;
    }
  }

  // DGRAD only: does this block need mirror image `img` (uniform across the block)?
  static VBX_HD bool dgrad_need_img(const GemmP& P, Blk blk, int img) {
    return dgrad_tile_needs_img(P, blk.z, blk.x * TN, TN, img);
  }

  // DGRAD only: set up the loaders for one mirror image; returns the reduction length
  static VBX_HD int dgrad_begin_img(const GemmP& P, int tid, TS& s, int img) {
    s.img = img;
    s.im = dgrad_img(P, s.ph.r, img);
    s.Kred = P.Cout_g * s.im.ntaps;
    int n = s.n_base + s.b_nl;
    s.b_valid = n < s.N;
    if (s.b_valid) {
      int b = n / s.ph.Up, u = s.ph.r + (n % s.ph.Up) * P.stride;
      int p;
      if (img == 0) p = u;
      else if (img == 1) { p = -u; if (u < 1 || u > P.refl) s.b_valid = false; }
      else { p = 2 * (P.Tin - 1) - u; if (u > P.Tin - 2 || u < P.Tin - 1 - P.refl) s.b_valid = false; }
      s.b_base = ((long long)b * P.Cout + s.g * P.Cout_g) * P.Tout;
      s.b_t = (p + P.pad) / P.stride - s.im.kd0_hi;     // p + pad >= 0 because refl <= pad
    }
    if (s.im.ntaps > 0) { s.b_c = s.b_r0 / s.im.ntaps; s.b_k = s.b_r0 % s.im.ntaps; }
    else { s.b_c = 0; s.b_k = 0; }
    return s.Kred;
  }

  static VBX_HD int num_chunks(const TS& s) { return (s.Kred + KC - 1) / KC; }

  // ---------------------------------------------------------------- global -> smem
  static VBX_HD void load_chunk(const GemmP& P, int tid, TS& s, int chunk, float* As, float* Bs) {
    // ---- A tile: As[kkl][m], lanes walk kk (contiguous in memory for every mode)
    {
      int kk = chunk * KC + s.a_kkl;
      bool kv = kk < s.Kred;
      long long koff = 0;
      long long mstride = 0;
      if (MODE == FWD) { koff = kk; mstride = s.Kred; }
      else if (MODE == SCATTER) { koff = kk; mstride = P.Cout_g; }
      else if (MODE == DGRAD) {
        int co = 0, j = 0;
        if (kv) { co = kk / s.im.ntaps; j = kk % s.im.ntaps; }
        koff = (long long)co * P.K + s.im.k0 + j * s.im.kstep;
        mstride = (long long)P.Cout_g * P.K;
      } else {  // WGRAD: A = dy[b][g*Cout_g + m][t]
        int r = s.red_lo + kk;
        int b = 0, t = 0;
        if (kv) { b = r / P.Tout; t = r % P.Tout; }
        koff = ((long long)b * P.Cout + s.g * P.Cout_g + s.m_base) * P.Tout + t;
        mstride = P.Tout;
      }
      const float* src = MODE == WGRAD ? P.DY : P.W;
      long long base = MODE == WGRAD ? 0 : s.a_base;
      for (int ml = s.a_m0; ml < TM; ml += C::AROWS) {
        float v = 0.f;
        if (kv && s.m_base + ml < s.M) v = src[base + (long long)ml * mstride + koff];
        As[s.a_kkl * C::LDA + ml] = v;
      }
    }
    // ---- B tile: Bs[kkl][n]
    if (MODE == FWD && !BK) {
      for (int kkl = s.b_r0; kkl < KC; kkl += C::BROWS_N) {
        float v = 0.f;
        if (s.b_valid && s.b_c < P.Cin_g) {
          int p = map_pos(s.b_t + s.b_k * P.dil, P.Tin, P.refl);
          if (p >= 0) v = P.X[s.b_base + (long long)s.b_c * P.Tin + p];
        }
        Bs[kkl * C::LDB + s.b_nl] = v;
        s.b_k += C::BROWS_N;
        while (s.b_k >= P.K) { s.b_k -= P.K; ++s.b_c; }
      }
    } else if (MODE == FWD && BK) {
      bool kv = s.bk_c < P.Cin_g;
      VBX_UNROLL
      for (int c = 0; c < NBCOL; ++c) {
        float v = 0.f;
        if (kv && s.bk_t[c] != INVALID) {
          int p = map_pos(s.bk_t[c] + s.bk_k * P.dil, P.Tin, P.refl);
          if (p >= 0) v = P.X[s.bk_off[c] + (long long)s.bk_c * P.Tin + p];
        }
        Bs[s.bk_kkl * C::LDB + s.bk_n0 + c * (NT / KC)] = v;
      }
      s.bk_k += KC;
      if (s.bk_k >= P.K) { s.bk_c += s.bk_k / P.K; s.bk_k %= P.K; }
    } else if (MODE == DGRAD) {
      for (int kkl = s.b_r0; kkl < KC; kkl += C::BROWS_N) {
        float v = 0.f;
        if (s.b_valid && s.b_c < P.Cout_g && s.im.ntaps > 0) {
          int t = s.b_t - s.b_k * s.im.tstep;
          if (t >= 0 && t < P.Tout) v = P.X[s.b_base + (long long)s.b_c * P.Tout + t];
        }
        Bs[kkl * C::LDB + s.b_nl] = v;
        if (s.im.ntaps > 0) {
          s.b_k += C::BROWS_N;
          while (s.b_k >= s.im.ntaps) { s.b_k -= s.im.ntaps; ++s.b_c; }
        }
      }
    } else if (MODE == WGRAD) {
      int kk = chunk * KC + s.bk_kkl;
      bool kv = kk < s.Kred;
      int r = s.red_lo + kk;
      int b = 0, t = 0;
      if (kv) { b = r / P.Tout; t = r % P.Tout; }
      long long xb = (long long)b * P.Cin * P.Tin;
      int ts = t * P.stride;
      VBX_UNROLL
      for (int c = 0; c < NBCOL; ++c) {
        float v = 0.f;
        if (kv && s.bk_t[c] != INVALID) {
          int p = map_pos(ts + s.bk_t[c], P.Tin, P.refl);
          if (p >= 0) v = P.X[xb + s.bk_off[c] + p];
        }
        Bs[s.bk_kkl * C::LDB + s.bk_n0 + c * (NT / KC)] = v;
      }
    } else {  // SCATTER: B = dy[b][g*Cout_g + co][t], rows are co
      for (int kkl = s.b_r0; kkl < KC; kkl += C::BROWS_N) {
        float v = 0.f;
        int co = chunk * KC + kkl;
        if (s.b_valid && co < P.Cout_g) v = P.X[s.b_base + (long long)co * P.Tout];
        Bs[kkl * C::LDB + s.b_nl] = v;
      }
    }
  }

  // ---------------------------------------------------------------- smem -> FMA
  static VBX_HD int row_of(int ty, int i) { return (i / C::RMV) * (TM / C::RMG) + ty * C::RMV + (i % C::RMV); }
  static VBX_HD int col_of(int tx, int j) { return (j / C::RNV) * (TN / C::RNG) + tx * C::RNV + (j % C::RNV); }

  static VBX_HD void compute_chunk(int tid, TS& s, const float* As, const float* Bs) {
    const int ty = tid / C::TX, tx = tid % C::TX;
VBX_UNROLL
    for (int kk = 0; kk < KC; ++kk) {
      float a[RM], b[RN];
VBX_UNROLL
      for (int ig = 0; ig < C::RMG; ++ig) {
        const float* ap = As + kk * C::LDA + ig * (TM / C::RMG) + ty * C::RMV;
        if (C::RMV == 4) {
          float4 v = *reinterpret_cast<const float4*>(ap);
          a[ig * 4 + 0] = v.x; a[ig * 4 + 1] = v.y; a[ig * 4 + 2] = v.z; a[ig * 4 + 3] = v.w;
        } else {
          VBX_UNROLL
          for (int i = 0; i < C::RMV; ++i) a[ig * C::RMV + i] = ap[i];
        }
      }
VBX_UNROLL
      for (int jg = 0; jg < C::RNG; ++jg) {
        float4 v = *reinterpret_cast<const float4*>(Bs + kk * C::LDB + jg * (TN / C::RNG) + tx * 4);
        b[jg * 4 + 0] = v.x; b[jg * 4 + 1] = v.y; b[jg * 4 + 2] = v.z; b[jg * 4 + 3] = v.w;
      }
VBX_UNROLL
      for (int i = 0; i < RM; ++i)
VBX_UNROLL
        for (int j = 0; j < RN; ++j) s.acc[i][j] = fmaf(a[i], b[j], s.acc[i][j]);
    }
  }

  // ---------------------------------------------------------------- epilogue
  static VBX_HD float finish(const GemmP& P, float v, int ch, long long idx) {
    if (P.bias) v += P.bias[ch];
    if (P.mask) P.mask[idx] = v > 0.f ? 1 : 0;
    if (P.slope != 1.f) v = v > 0.f ? v : v * P.slope;
    if (P.res) v += P.res[idx];
    if (P.gate) {
      const bool fm = P.fm_other != nullptr;
      v = gate_apply(v, P.gate[idx], fm, fm ? P.fm_other[idx] : 0.f, fm ? P.fm_coef[0] : 0.f, fm ? P.fm_coef[1] : 0.f,
                     P.gate_slope);
    }
    if (P.beta != 0.f) v += P.beta * P.Y[idx];
    return v;
  }

  template <class AtomicAdd>
  static VBX_HD void epilogue(const GemmP& P, Blk blk, int tid, TS& s, AtomicAdd atomic_add) {
    const int ty = tid / C::TX, tx = tid % C::TX;
    VBX_UNROLL
    for (int jg = 0; jg < C::RNG; ++jg) {
      const int n0 = s.n_base + jg * (TN / C::RNG) + tx * 4;
      if (n0 >= s.N) continue;
      if (MODE == FWD || MODE == DGRAD) {
        // decode the 4 columns once
        long long cbase[4]; bool cv[4];
        const int Tlen = MODE == FWD ? P.Tout : P.Tin;
        const int Ctot = MODE == FWD ? P.Cout : P.Cin;
        const int Cg = MODE == FWD ? P.Cout_g : P.Cin_g;
        VBX_UNROLL
        for (int j = 0; j < 4; ++j) {
          int n = n0 + j;
          cv[j] = n < s.N;
          int b, t;
          if (MODE == FWD) { b = cv[j] ? n / P.Tout : 0; t = cv[j] ? n % P.Tout : 0; }
          else { b = cv[j] ? n / s.ph.Up : 0; t = s.ph.r + (cv[j] ? n % s.ph.Up : 0) * P.stride; }
          cbase[j] = ((long long)b * Ctot + s.g * Cg) * Tlen + t;
        }
        const bool vec = MODE == FWD && cv[3] && (cbase[3] - cbase[0] == 3) &&
                         ((cbase[0] & 3) == 0) && ((Tlen & 3) == 0) && !P.mask;
        VBX_UNROLL
        for (int i = 0; i < RM; ++i) {
          int m = s.m_base + row_of(ty, i);
          if (m >= s.M) continue;
          int ch = s.g * Cg + m;
          long long roff = (long long)m * Tlen;
          if (vec) {
            float4 o;
            o.x = finish(P, s.acc[i][jg * 4 + 0], ch, cbase[0] + roff + 0);
            o.y = finish(P, s.acc[i][jg * 4 + 1], ch, cbase[0] + roff + 1);
            o.z = finish(P, s.acc[i][jg * 4 + 2], ch, cbase[0] + roff + 2);
            o.w = finish(P, s.acc[i][jg * 4 + 3], ch, cbase[0] + roff + 3);
            *reinterpret_cast<float4*>(P.Y + cbase[0] + roff) = o;
          } else {
            VBX_UNROLL
            for (int j = 0; j < 4; ++j)
              if (cv[j]) {
                long long idx = cbase[j] + roff;
                P.Y[idx] = finish(P, s.acc[i][jg * 4 + j], ch, idx);
              }
          }
        }
      } else if (MODE == WGRAD) {
        VBX_UNROLL
        for (int i = 0; i < RM; ++i) {
          int m = s.m_base + row_of(ty, i);
          if (m >= s.M) continue;
          long long row = (long long)(s.g * P.Cout_g + m) * s.N;
          VBX_UNROLL
          for (int j = 0; j < 4; ++j)
            if (n0 + j < s.N) atomic_add(P.Y + row + n0 + j, s.acc[i][jg * 4 + j]);
        }
      } else {  // SCATTER
        VBX_UNROLL
        for (int j = 0; j < 4; ++j) {
          int n = n0 + j;
          if (n >= s.N) continue;
          int b = n / P.Tout, t = n % P.Tout;
          long long xb = ((long long)b * P.Cin + s.g * P.Cin_g) * P.Tin;
          VBX_UNROLL
          for (int i = 0; i < RM; ++i) {
            int m = s.m_base + row_of(ty, i);
            if (m >= s.M) continue;
            int ci = m / P.K, k = m % P.K;
            int p = map_pos(t * P.stride + k * P.dil - P.pad, P.Tin, P.refl);
            if (p >= 0) atomic_add(P.Y + xb + (long long)ci * P.Tin + p, s.acc[i][jg * 4 + j]);
          }
        }
      }
    }
  }
};

}  // namespace vbx
