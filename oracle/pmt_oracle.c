/*
 * pmt_oracle.c -- CPU restatement of the reference's stereo cost-volume hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (pmt_learning_for_semantic_segmentation_and_disparity_b200/) never imports it.
 *
 * Parity status of each function:
 *   corr_*      : PARITY UNPINNED by the reference -- the arithmetic lives in the un-vendored,
 *                 un-pinned PyPI package `spatial-correlation-sampler` (reference
 *                 scripts/scriptsDocker/Torch/Dockerfile:53, README.md:7).  This file restates
 *                 that package's published CPU algorithm (upstream correlation.cpp:
 *                 correlate_patch / correlate_patch_grad loop nests, fp32 accumulation in
 *                 channel order, output zero-initialised, NOT divided by C) and is anchored on
 *                 the reference's call sites: models/dsnet_t2.py:129-133,221-223,841-851,879-884,
 *                 models/torch_dsnet.py:133-138, models/dsnet_t2_warp.py:615-619,664.
 *   concat_*    : pinned against the reference's own loop (models_psmnet/stackhourglass.py:110-119)
 *                 and matchshifted (models_psmnet/submodule.py:45-54) run in the build container;
 *                 fixtures in tests/golden/ (oracle/make_golden.py).
 *   dispreg_*, softargmin_* : pinned against models_psmnet/submodule.py:56-64 +
 *                 F.softmax (stackhourglass.py:151,155); fixtures in tests/golden/.
 *   warp_*      : pinned against models/torch_dsnet.py:10-86 (apply_disparity); fixtures in
 *                 tests/golden/.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 * All tensors are dense, row-major ("contiguous" in torch terms), fp32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX4(n, c, h, w, C, H, W) ((((int64_t)(n) * (C) + (c)) * (H) + (h)) * (int64_t)(W) + (w))

int pmt_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void pmt_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * a1. SpatialCorrelationSampler (general form of the upstream CPU implementation).
 *   out[n,ph,pw,h,w] = sum_c sum_{i<kH, j<kW} in1[n,c,i1,j1] * in2[n,c,i1+sh,j1+sw]
 *   i1 = -padH + h*dH + i*dilH,  sh = (ph - (patchH-1)/2) * dilpH      (integer division)
 *   terms with either index outside the image are skipped; no normalisation by C.
 * Call sites: models/dsnet_t2.py:129-133 (2-D 17x17), :841-851 (1x17), models/torch_dsnet.py:133-138
 * (1x21, dilation_patch 4).  Output shape (B, patchH, patchW, oH, oW), see dsnet_t2.py:222,881.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int kH, kW, patchH, patchW, padH, padW, dilH, dilW, dilpH, dilpW, dH, dW;
} pmt_corr_params;

static void corr_out_size(int iH, int iW, const pmt_corr_params* p, int* oH, int* oW) {
  int dkH = (p->kH - 1) * p->dilH + 1;
  int dkW = (p->kW - 1) * p->dilW + 1;
  *oH = (iH + 2 * p->padH - dkH) / p->dH + 1;
  *oW = (iW + 2 * p->padW - dkW) / p->dW + 1;
}

void pmt_oracle_corr_out_size(int iH, int iW, const int* params12, int* oH, int* oW) {
  pmt_corr_params p;
  memcpy(&p, params12, sizeof(p));
  corr_out_size(iH, iW, &p, oH, oW);
}

void pmt_oracle_corr_fwd(const float* in1, const float* in2, float* out, int B, int C, int iH,
                         int iW, const int* params12) {
  pmt_corr_params p;
  memcpy(&p, params12, sizeof(p));
  int oH, oW;
  corr_out_size(iH, iW, &p, &oH, &oW);
  const int radH = (p.patchH - 1) / 2, radW = (p.patchW - 1) / 2;
  const int64_t plane = (int64_t)iH * iW;
  /* Upstream accumulates `*dst += v1*v2` with the channel loop outermost inside correlate_patch.
   * We keep exactly that per-element order (c, then i, then j) but walk a whole output row per
   * (c,i,j) so the memory accesses are contiguous; every output element sees the same sequence of
   * fp32 additions as in the upstream loop nest. */
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < B; ++n) {
    for (int ph = 0; ph < p.patchH; ++ph) {
      const float* a = in1 + (int64_t)n * C * plane;
      const float* b = in2 + (int64_t)n * C * plane;
      for (int pw = 0; pw < p.patchW; ++pw) {
        const int sh = (ph - radH) * p.dilpH, sw = (pw - radW) * p.dilpW;
        for (int h = 0; h < oH; ++h) {
          float* orow = out + ((((int64_t)n * p.patchH + ph) * p.patchW + pw) * oH + h) * (int64_t)oW;
          for (int w = 0; w < oW; ++w) orow[w] = 0.0f; /* output is zero-initialised */
          const int u = -p.padH + h * p.dH;
          for (int c = 0; c < C; ++c) {
            for (int i = 0; i < p.kH; ++i) {
              const int i1 = u + i * p.dilH, i2 = i1 + sh;
              if (i1 < 0 || i1 >= iH || i2 < 0 || i2 >= iH) continue;
              const float* arow = a + c * plane + (int64_t)i1 * iW;
              const float* brow = b + c * plane + (int64_t)i2 * iW;
              for (int j = 0; j < p.kW; ++j) {
                for (int w = 0; w < oW; ++w) {
                  const int j1 = -p.padW + w * p.dW + j * p.dilW, j2 = j1 + sw;
                  if (j1 < 0 || j1 >= iW || j2 < 0 || j2 >= iW) continue;
                  orow[w] += arow[j1] * brow[j2];
                }
              }
            }
          }
        }
      }
    }
  }
}

/* Backward: same loop nest as upstream correlate_patch_grad -- for every output element the two
 * gradients are scattered in (ph, pw, h, w) order; one OpenMP thread per batch item so the
 * accumulation order inside a batch item is exactly sequential. */
void pmt_oracle_corr_bwd(const float* in1, const float* in2, const float* gout, float* g1,
                         float* g2, int B, int C, int iH, int iW, const int* params12) {
  pmt_corr_params p;
  memcpy(&p, params12, sizeof(p));
  int oH, oW;
  corr_out_size(iH, iW, &p, &oH, &oW);
  const int radH = (p.patchH - 1) / 2, radW = (p.patchW - 1) / 2;
  const int64_t plane = (int64_t)iH * iW;
  memset(g1, 0, sizeof(float) * (size_t)B * C * plane);
  memset(g2, 0, sizeof(float) * (size_t)B * C * plane);
  /* parallel over (n, c): every (n, c) plane is independent and keeps the (ph,pw,h,w) order */
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < B; ++n) {
    for (int c = 0; c < C; ++c) {
      const float* a = in1 + ((int64_t)n * C + c) * plane;
      const float* b = in2 + ((int64_t)n * C + c) * plane;
      float* ga = g1 + ((int64_t)n * C + c) * plane;
      float* gb = g2 + ((int64_t)n * C + c) * plane;
      for (int ph = 0; ph < p.patchH; ++ph) {
        for (int pw = 0; pw < p.patchW; ++pw) {
          const int sh = (ph - radH) * p.dilpH, sw = (pw - radW) * p.dilpW;
          const float* g =
              gout + (((int64_t)n * p.patchH + ph) * p.patchW + pw) * (int64_t)oH * oW;
          for (int h = 0; h < oH; ++h) {
            for (int w = 0; w < oW; ++w) {
              const float go = g[(int64_t)h * oW + w];
              const int u = -p.padH + h * p.dH, v = -p.padW + w * p.dW;
              for (int i = 0; i < p.kH; ++i) {
                const int i1 = u + i * p.dilH, i2 = i1 + sh;
                if (i1 < 0 || i1 >= iH || i2 < 0 || i2 >= iH) continue;
                for (int j = 0; j < p.kW; ++j) {
                  const int j1 = v + j * p.dilW, j2 = j1 + sw;
                  if (j1 < 0 || j1 >= iW || j2 < 0 || j2 >= iW) continue;
                  const float v1 = a[(int64_t)i1 * iW + j1], v2 = b[(int64_t)i2 * iW + j2];
                  gb[(int64_t)i2 * iW + j2] += go * v1;
                  ga[(int64_t)i1 * iW + j1] += go * v2;
                }
              }
            }
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * a2. PSMNet concat cost volume (models_psmnet/stackhourglass.py:110-119):
 *   cost[b, c,   i, h, w] = ref[b,c,h,w]     if w >= i else 0
 *   cost[b, C+c, i, h, w] = tgt[b,c,h,w-i]   if w >= i else 0        i in [0, D)
 * matchshifted(left,right,shift) (submodule.py:45-54) is the i == shift slice.
 * ------------------------------------------------------------------------------------------ */
void pmt_oracle_concat_fwd(const float* ref, const float* tgt, float* cost, int B, int C, int D,
                           int H, int W) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int c2 = 0; c2 < 2 * C; ++c2) {
      const int is_tgt = c2 >= C;
      const float* src = (is_tgt ? tgt : ref) + IDX4(b, is_tgt ? c2 - C : c2, 0, 0, C, H, W);
      for (int i = 0; i < D; ++i) {
        float* dst = cost + ((((int64_t)b * 2 * C + c2) * D + i) * H) * (int64_t)W;
        for (int h = 0; h < H; ++h) {
          for (int w = 0; w < W; ++w) {
            float v = 0.0f;
            if (w >= i) v = is_tgt ? src[(int64_t)h * W + (w - i)] : src[(int64_t)h * W + w];
            dst[(int64_t)h * W + w] = v;
          }
        }
      }
    }
  }
}

/* Autograd of the slice assignments: g_ref[w] = sum_{i<=w} g[c,i,h,w];
 * g_tgt[w'] = sum_{i: w'+i<W} g[C+c,i,h,w'+i].  Autograd accumulates the slices in reverse
 * order of the forward loop (i = D-1 .. 0); we keep that order. */
void pmt_oracle_concat_bwd(const float* gcost, float* gref, float* gtgt, int B, int C, int D, int H,
                           int W) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int c2 = 0; c2 < 2 * C; ++c2) {
      const int is_tgt = c2 >= C;
      float* dst = (is_tgt ? gtgt : gref) + IDX4(b, is_tgt ? c2 - C : c2, 0, 0, C, H, W);
      for (int64_t k = 0; k < (int64_t)H * W; ++k) dst[k] = 0.0f;
      for (int i = D - 1; i >= 0; --i) {
        const float* g = gcost + ((((int64_t)b * 2 * C + c2) * D + i) * H) * (int64_t)W;
        for (int h = 0; h < H; ++h) {
          for (int w = i; w < W; ++w) {
            if (is_tgt)
              dst[(int64_t)h * W + (w - i)] += g[(int64_t)h * W + w];
            else
              dst[(int64_t)h * W + w] += g[(int64_t)h * W + w];
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * a3. disparityregression (models_psmnet/submodule.py:56-64): out[b,h,w] = sum_d x[b,d,h,w]*d
 * (torch.sum over dim 1 of x*disp; fp32 sequential order over d is what we restate) and the
 * fused F.softmax(dim=1) + disparityregression pair (stackhourglass.py:151,155).
 * ------------------------------------------------------------------------------------------ */
void pmt_oracle_dispreg_fwd(const float* x, float* out, int B, int D, int H, int W) {
  const int64_t plane = (int64_t)H * W;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)B * plane; ++q) {
    const int64_t b = q / plane, k = q % plane;
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) acc += x[(b * D + d) * plane + k] * (float)d;
    out[q] = acc;
  }
}

void pmt_oracle_dispreg_bwd(const float* gout, float* gx, int B, int D, int H, int W) {
  const int64_t plane = (int64_t)H * W;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)B * plane; ++q) {
    const int64_t b = q / plane, k = q % plane;
    for (int d = 0; d < D; ++d) gx[(b * D + d) * plane + k] = (float)d * gout[q];
  }
}

void pmt_oracle_softargmin_fwd(const float* cost, float* out, int B, int D, int H, int W) {
  const int64_t plane = (int64_t)H * W;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)B * plane; ++q) {
    const int64_t b = q / plane, k = q % plane;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, cost[(b * D + d) * plane + k]);
    float s = 0.0f;
    for (int d = 0; d < D; ++d) s += expf(cost[(b * D + d) * plane + k] - m);
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) acc += (expf(cost[(b * D + d) * plane + k] - m) / s) * (float)d;
    out[q] = acc;
  }
}

/* g_cost[b,d,h,w] = g[b,h,w] * p_d * (d - out[b,h,w])   (softmax backward through the regression) */
void pmt_oracle_softargmin_bwd(const float* cost, const float* gout, float* gcost, int B, int D,
                               int H, int W) {
  const int64_t plane = (int64_t)H * W;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)B * plane; ++q) {
    const int64_t b = q / plane, k = q % plane;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) m = fmaxf(m, cost[(b * D + d) * plane + k]);
    double s = 0.0, e1 = 0.0;
    for (int d = 0; d < D; ++d) {
      double e = exp((double)cost[(b * D + d) * plane + k] - (double)m);
      s += e;
      e1 += e * d;
    }
    const double o = e1 / s;
    for (int d = 0; d < D; ++d) {
      double pd = exp((double)cost[(b * D + d) * plane + k] - (double)m) / s;
      gcost[(b * D + d) * plane + k] = (float)((double)gout[q] * pd * ((double)d - o));
    }
  }
}

/* f1. F.upsample(cost3, [D,H,W], mode='trilinear') -> squeeze -> softmax(dim=1) -> disparityregression,
 * models_psmnet/stackhourglass.py:149-155.  ATen's align_corners=False rule: scale = in/out (float),
 * src = scale*(dst+0.5)-0.5 clamped at 0, i0=(int)src, i1=i0+(i0<in-1), l1=src-i0, l0=1-l1, and the trilinear blend
 * t0*(h0*(w0*v000+w1*v001)+h1*(w0*v010+w1*v011)) + t1*(...).  Pinned by tests/golden/upsoftargmin_small.npz. */
static void up_src(float scale, int dst, int in, int* i0, int* i1, float* l0, float* l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.0f) src = 0.0f;
  *i0 = (int)src;
  if (*i0 > in - 1) *i0 = in - 1;
  *i1 = *i0 + ((*i0 < in - 1) ? 1 : 0);
  *l1 = src - (float)*i0;
  *l0 = 1.0f - *l1;
}

void pmt_oracle_upsample_softargmin_fwd(const float* low, float* out, int B, int Dq, int Hq, int Wq, int D, int H,
                                        int W) {
  const float sd = (float)Dq / (float)D, sh = (float)Hq / (float)H, sw = (float)Wq / (float)W;
  const int64_t qplane = (int64_t)Hq * Wq;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < (int64_t)B * H * W; ++q) {
    const int w = (int)(q % W), h = (int)((q / W) % H), b = (int)(q / ((int64_t)W * H));
    int h0, h1, w0, w1;
    float hl0, hl1, wl0, wl1;
    up_src(sh, h, Hq, &h0, &h1, &hl0, &hl1);
    up_src(sw, w, Wq, &w0, &w1, &wl0, &wl1);
    const float* base = low + (int64_t)b * Dq * qplane;
    float* c = (float*)malloc(sizeof(float) * (size_t)D);
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) {
      int t0, t1;
      float tl0, tl1;
      up_src(sd, d, Dq, &t0, &t1, &tl0, &tl1);
      const float* p0 = base + (int64_t)t0 * qplane;
      const float* p1 = base + (int64_t)t1 * qplane;
      const float a0 = hl0 * (wl0 * p0[h0 * Wq + w0] + wl1 * p0[h0 * Wq + w1]) +
                       hl1 * (wl0 * p0[h1 * Wq + w0] + wl1 * p0[h1 * Wq + w1]);
      const float a1 = hl0 * (wl0 * p1[h0 * Wq + w0] + wl1 * p1[h0 * Wq + w1]) +
                       hl1 * (wl0 * p1[h1 * Wq + w0] + wl1 * p1[h1 * Wq + w1]);
      c[d] = tl0 * a0 + tl1 * a1;
      m = fmaxf(m, c[d]);
    }
    float ssum = 0.0f, acc = 0.0f;
    for (int d = 0; d < D; ++d) ssum += expf(c[d] - m);
    for (int d = 0; d < D; ++d) acc += (expf(c[d] - m) / ssum) * (float)d;
    out[q] = acc;
    free(c);
  }
}

/* f1 backward: gradient of sum(gout * pred) w.r.t. the low-res logits, by the chain rule through the same steps
 * (autograd of the reference sequence): g_x[d] = gout * p_d * (d - pred), scattered onto the 8 trilinear taps.
 * Serial scatter in double (the oracle is not the thing measured).  Pinned by gcost3 of upsoftargmin_small.npz. */
void pmt_oracle_upsample_softargmin_bwd(const float* low, const float* gout, float* glow, int B, int Dq, int Hq, int Wq,
                                        int D, int H, int W) {
  const float sd = (float)Dq / (float)D, sh = (float)Hq / (float)H, sw = (float)Wq / (float)W;
  const int64_t qplane = (int64_t)Hq * Wq, nlow = (int64_t)B * Dq * qplane;
  double* acc = (double*)calloc((size_t)nlow, sizeof(double));
  float* c = (float*)malloc(sizeof(float) * (size_t)D);
  for (int64_t q = 0; q < (int64_t)B * H * W; ++q) {
    const int w = (int)(q % W), h = (int)((q / W) % H), b = (int)(q / ((int64_t)W * H));
    int h0, h1, w0, w1;
    float hl0, hl1, wl0, wl1;
    up_src(sh, h, Hq, &h0, &h1, &hl0, &hl1);
    up_src(sw, w, Wq, &w0, &w1, &wl0, &wl1);
    const float* base = low + (int64_t)b * Dq * qplane;
    float m = -INFINITY;
    for (int d = 0; d < D; ++d) {
      int t0, t1;
      float tl0, tl1;
      up_src(sd, d, Dq, &t0, &t1, &tl0, &tl1);
      const float* p0 = base + (int64_t)t0 * qplane;
      const float* p1 = base + (int64_t)t1 * qplane;
      const float a0 = hl0 * (wl0 * p0[h0 * Wq + w0] + wl1 * p0[h0 * Wq + w1]) +
                       hl1 * (wl0 * p0[h1 * Wq + w0] + wl1 * p0[h1 * Wq + w1]);
      const float a1 = hl0 * (wl0 * p1[h0 * Wq + w0] + wl1 * p1[h0 * Wq + w1]) +
                       hl1 * (wl0 * p1[h1 * Wq + w0] + wl1 * p1[h1 * Wq + w1]);
      c[d] = tl0 * a0 + tl1 * a1;
      m = fmaxf(m, c[d]);
    }
    double ssum = 0.0, pred = 0.0;
    for (int d = 0; d < D; ++d) ssum += exp((double)c[d] - (double)m);
    for (int d = 0; d < D; ++d) pred += exp((double)c[d] - (double)m) / ssum * (double)d;
    double* ab = acc + (int64_t)b * Dq * qplane;
    for (int d = 0; d < D; ++d) {
      int t0, t1;
      float tl0, tl1;
      up_src(sd, d, Dq, &t0, &t1, &tl0, &tl1);
      const double gx = (double)gout[q] * (exp((double)c[d] - (double)m) / ssum) * ((double)d - pred);
      const int ts[2] = {t0, t1};
      const double tw[2] = {tl0, tl1};
      const int hs[2] = {h0, h1};
      const double hw[2] = {hl0, hl1};
      const int ws[2] = {w0, w1};
      const double ww[2] = {wl0, wl1};
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
          for (int k = 0; k < 2; ++k) ab[(int64_t)ts[i] * qplane + hs[j] * Wq + ws[k]] += gx * tw[i] * hw[j] * ww[k];
    }
  }
  for (int64_t i = 0; i < nlow; ++i) glow[i] = (float)acc[i];
  free(acc);
  free(c);
}

/* ------------------------------------------------------------------------------------------
 * a4. apply_disparity(img, x_offset, wrap_mode='edge') -- models/torch_dsnet.py:10-86.
 * Every arithmetic step is kept in fp32 exactly as the reference does it, INCLUDING the flat
 * gather index built in float32 (torch_dsnet.py:59-70): idx = fl(fl(base + fl(y0*dim2)) + x0).
 *   x  = clamp(fl(w + off), 0, W-1); x0 = floor(x); x1 = min(x0+1, W-1)
 *   out = fl(fl((x1-x)*pix_l) + fl((x-x0)*pix_r))
 * ------------------------------------------------------------------------------------------ */
static inline int64_t warp_flat_index(int n, int h, float xf, int H, int W) {
  /* base = dim1 * arange(N) in fp32 (torch_dsnet.py:59), base_y0 = base + y0*dim2 (:65) */
  volatile float base = (float)((int64_t)W * H) * (float)n;
  volatile float y0w = (float)h * (float)W;
  volatile float by = base + y0w;
  volatile float idx = by + xf;
  return (int64_t)idx;
}

void pmt_oracle_warp_fwd(const float* img, const float* off, float* out, int N, int C, int H, int W) {
  const int64_t plane = (int64_t)H * W, total = (int64_t)N * plane;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < total; ++q) {
    const int n = (int)(q / plane), h = (int)((q % plane) / W), w = (int)(q % W);
    volatile float x = (float)w + off[q];
    float xc = fminf(fmaxf(x, 0.0f), (float)(W - 1));
    float x0 = floorf(xc);
    float x1 = fminf(x0 + 1.0f, (float)(W - 1));
    int64_t il = warp_flat_index(n, h, x0, H, W), ir = warp_flat_index(n, h, x1, H, W);
    if (il > total - 1) il = total - 1; /* the reference would raise here; keep memory-safe */
    if (ir > total - 1) ir = total - 1;
    volatile float wl = x1 - xc, wr = xc - x0;
    for (int c = 0; c < C; ++c) {
      /* im_flat = img.permute(1,0,2,3).view(C, N*H*W): element (c, idx) */
      const float pl = img[IDX4(il / plane, c, 0, 0, C, H, W) + il % plane];
      const float pr = img[IDX4(ir / plane, c, 0, 0, C, H, W) + ir % plane];
      volatile float a = wl * pl, b2 = wr * pr;
      out[IDX4(n, c, h, w, C, H, W)] = a + b2;
    }
  }
}

/* Backward as autograd derives it from the reference graph: gather backward scatters
 * wl*g / wr*g into the image (scatter_add), the offset receives sum_c g*(pix_r - pix_l) where the
 * clamp is not saturated (closed interval 0 <= w+off <= W-1). One thread per (n) so the
 * scatter order is sequential and deterministic. */
void pmt_oracle_warp_bwd(const float* img, const float* off, const float* gout, float* gimg,
                         float* goff, int N, int C, int H, int W) {
  const int64_t plane = (int64_t)H * W, total = (int64_t)N * plane;
  memset(gimg, 0, sizeof(float) * (size_t)N * C * plane);
  /* sequential: faithful fp32 indices may cross batch items */
  for (int64_t q = 0; q < total; ++q) {
    const int n = (int)(q / plane), h = (int)((q % plane) / W), w = (int)(q % W);
    volatile float x = (float)w + off[q];
    float xc = fminf(fmaxf(x, 0.0f), (float)(W - 1));
    float x0 = floorf(xc);
    float x1 = fminf(x0 + 1.0f, (float)(W - 1));
    int64_t il = warp_flat_index(n, h, x0, H, W), ir = warp_flat_index(n, h, x1, H, W);
    if (il > total - 1) il = total - 1;
    if (ir > total - 1) ir = total - 1;
    const float wl = x1 - xc, wr = xc - x0;
    float gl = 0.0f, gr = 0.0f;
    for (int c = 0; c < C; ++c) {
      const int64_t al = IDX4(il / plane, c, 0, 0, C, H, W) + il % plane;
      const int64_t ar = IDX4(ir / plane, c, 0, 0, C, H, W) + ir % plane;
      const float g = gout[IDX4(n, c, h, w, C, H, W)];
      gimg[al] += wl * g;
      gimg[ar] += wr * g;
      gl += g * img[al];
      gr += g * img[ar];
    }
    const int pass = (x >= 0.0f) && (x <= (float)(W - 1));
    goff[q] = pass ? (gr - gl) : 0.0f;
  }
}
