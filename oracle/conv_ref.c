/* conv_ref.c — plain-C restatement of the convolution primitives the reference reaches through torch
 * (un-vendored third-party dependency, unpinned in the reference's requirements.txt:1; PyTorch 2.11.0 here).
 * TEST INFRASTRUCTURE: only tests/ may load this; it pins what F.conv1d / F.conv_transpose1d /
 * F.avg_pool1d *mean* independently of torch, in double-accumulated direct loops.
 *
 * Call sites restated: torch.nn.Conv1d (src/models.py:15-32,55-60,81,96,195-204),
 * torch.nn.ConvTranspose1d (src/models.py:84-88), torch.nn.AvgPool1d(4,2,padding=2) (src/models.py:227-230).
 * Layouts are torch's: x [B][Cin][T], w [Cout][Cin/groups][K] (conv) or [Cin][Cout][K] (transpose).
 */
#include <stddef.h>

int ref_conv1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int T, int Cout,
               int K, int stride, int padding, int dilation, int groups) {
  if (Cin % groups || Cout % groups) return -1;
  const int Tout = (T + 2 * padding - dilation * (K - 1) - 1) / stride + 1;
  const int cig = Cin / groups, cog = Cout / groups;
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co) {
      const int g = co / cog;
      for (int t = 0; t < Tout; ++t) {
        double acc = bias ? bias[co] : 0.0;
        for (int ci = 0; ci < cig; ++ci)
          for (int k = 0; k < K; ++k) {
            const int ti = t * stride - padding + k * dilation;
            if (ti < 0 || ti >= T) continue;
            acc += (double)x[((size_t)b * Cin + g * cig + ci) * T + ti] *
                   (double)w[((size_t)co * cig + ci) * K + k];
          }
        y[((size_t)b * Cout + co) * Tout + t] = (float)acc;
      }
    }
  return Tout;
}

int ref_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int T,
                         int Cout, int K, int stride, int padding) {
  const int Tout = (T - 1) * stride - 2 * padding + K;
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co)
      for (int o = 0; o < Tout; ++o) {
        double acc = bias ? bias[co] : 0.0;
        for (int k = 0; k < K; ++k) {
          const int num = o + padding - k;
          if (num < 0 || num % stride) continue;
          const int i = num / stride;
          if (i >= T) continue;
          for (int ci = 0; ci < Cin; ++ci)
            acc += (double)x[((size_t)b * Cin + ci) * T + i] * (double)w[((size_t)ci * Cout + co) * K + k];
        }
        y[((size_t)b * Cout + co) * Tout + o] = (float)acc;
      }
  return Tout;
}

/* AvgPool1d, count_include_pad=True, ceil_mode=False */
int ref_avg_pool1d(const float* x, float* y, int BC, int T, int K, int stride, int padding) {
  const int Tout = (T + 2 * padding - K) / stride + 1;
  for (int r = 0; r < BC; ++r)
    for (int t = 0; t < Tout; ++t) {
      double acc = 0.0;
      for (int k = 0; k < K; ++k) {
        const int ti = t * stride - padding + k;
        if (ti >= 0 && ti < T) acc += x[(size_t)r * T + ti];
      }
      y[(size_t)r * Tout + t] = (float)(acc / K);
    }
  return Tout;
}

void ref_leaky_relu(float* x, size_t n, float slope) {
  for (size_t i = 0; i < n; ++i) x[i] = x[i] > 0.f ? x[i] : x[i] * slope;
}
