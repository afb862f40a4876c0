/* gamma.c -- discrete-Gamma category rates on the host.
 *
 * Replaces pll_compute_gamma_cats (reference src/gamma.c:221-284, mean method): the rates are
 * host-side inputs of the likelihood path (once per alpha proposal), exactly as in the reference.
 * The numerical recipes are the published ones the reference uses, restated here as structured
 * loops with the published constants:
 *   ln Gamma      Pike & Hill (1966) CACM Algorithm 291 (Stirling series)           gamma.c:96-130
 *   I(x, alpha)   Bhattacharjee (1970) Appl. Stat. AS 32 (series / continued fraction) :28-95
 *   z_p           Odeh & Evans (1974) Appl. Stat. AS 70                               :132-160
 *   chi2_p        Best & Roberts (1975) Appl. Stat. AS 91                             :162-219
 */
#include <math.h>
#include <stdlib.h>
#include "bpp_gpu_host.h"

static double ln_gamma(double alpha)
{
  double x = alpha, f = 0.0, z;
  if (x < 7.0)
  {
    f = 1.0;
    for (z = alpha; z < 7.0; z += 1.0) f *= z;     /* product alpha (alpha+1) ... below 7 */
    x = z;
    f = -log(f);
  }
  z = 1.0 / (x * x);
  return f + (x - 0.5) * log(x) - x + .918938533204673 +
         (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x;
}

static double incomplete_gamma(double x, double alpha, double ln_gamma_alpha)
{
  const double accurate = 1e-8, overflow = 1e30;
  double p = alpha, factor, gin, term, rn;
  if (x == 0) return 0;
  if (x < 0 || p <= 0) return -1;
  factor = exp(p * log(x) - x - ln_gamma_alpha);
  if (!(x > 1 && x >= p))
  {
    /* series expansion */
    gin = 1; term = 1; rn = p;
    do { rn += 1; term *= x / rn; gin += term; } while (term > accurate);
    return gin * factor / p;
  }
  /* continued fraction */
  {
    double a = 1 - p, b = a + x + 1, an, dif, pn[6];
    int i;
    term = 0;
    pn[0] = 1; pn[1] = x; pn[2] = x + 1; pn[3] = x * b;
    gin = pn[2] / pn[3];
    for (;;)
    {
      a += 1; b += 2; term += 1;
      an = a * term;
      for (i = 0; i < 2; ++i) pn[i + 4] = b * pn[i + 2] - an * pn[i];
      if (pn[5] != 0)
      {
        rn = pn[4] / pn[5];
        dif = fabs(gin - rn);
        if (dif <= accurate && dif <= accurate * rn) break;
        gin = rn;
      }
      for (i = 0; i < 4; ++i) pn[i] = pn[i + 2];
      if (fabs(pn[4]) >= overflow) for (i = 0; i < 4; ++i) pn[i] /= overflow;
    }
    return 1 - factor * gin;
  }
}

static double point_normal(double prob)
{
  const double a0 = -.322232431088, a1 = -1, a2 = -.342242088547, a3 = -.0204231210245, a4 = -.453642210148e-4;
  const double b0 = .0993484626060, b1 = .588581570495, b2 = .531103462366, b3 = .103537752850, b4 = .0038560700634;
  double p = prob, p1 = (p < 0.5 ? p : 1 - p), y, z;
  if (p1 < 1e-20) return -9999;
  y = sqrt(log(1 / (p1 * p1)));
  z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
  return p < 0.5 ? -z : z;
}

static double point_chi2(double prob, double v)
{
  const double e = .5e-6, aa = .6931471805;
  double p = prob, g, xx, c, ch, a, q, p1, p2, t, x, b, s1, s2, s3, s4, s5, s6;
  if (p < .000002 || p > .999998 || v <= 0) return -1;
  g = ln_gamma(v / 2);
  xx = v / 2; c = xx - 1;
  if (v < -1.24 * log(p))
  {
    ch = pow(p * xx * exp(g + xx * aa), 1 / xx);
    if (ch - e < 0) return ch;
  }
  else if (v <= .32)
  {
    ch = 0.4; a = log(1 - p);
    do
    {
      q = ch; p1 = 1 + ch * (4.67 + ch); p2 = ch * (6.73 + ch * (6.66 + ch));
      t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
      ch -= (1 - exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (fabs(q / ch - 1) - .01 > 0);
  }
  else
  {
    x = point_normal(p);
    p1 = 0.222222 / v; ch = v * pow(x * sqrt(p1) + 1 - p1, 3.0);
    if (ch > 2.2 * v + 6) ch = -2 * (log(1 - p) - c * log(.5 * ch) + g);
  }
  do
  {
    q = ch; p1 = .5 * ch;
    t = incomplete_gamma(p1, xx, g);
    if (t < 0.0) return -1;
    p2 = p - t;
    t = p2 * exp(xx * aa + g + p1 - c * log(ch));
    b = t / ch; a = 0.5 * t - b * c;
    s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
    s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
    s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
    s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
    s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
    s6 = (120 + c * (346 + 127 * c)) / 5040;
    ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (fabs(q / ch - 1) > e);
  return ch;
}

/* mean-of-category rates of the discrete Gamma(alpha, beta) with `categories` equal-probability
   classes (gamma.c:258-277); categories == 1 gives rate 1 */
int bppgpu_compute_gamma_cats(double alpha, double beta, unsigned int categories, double * output_rates)
{
  unsigned int i;
  const double mean = alpha / beta;
  double lnga1, * cut;
  if (categories == 0) return BPPGPU_FAILURE;
  if (categories == 1) { output_rates[0] = 1.0; return BPPGPU_SUCCESS; }
  cut = (double *)malloc(categories * sizeof(double));
  if (!cut) return BPPGPU_FAILURE;
  lnga1 = ln_gamma(alpha + 1);
  for (i = 0; i + 1 < categories; ++i) cut[i] = point_chi2((i + 1.0) / categories, 2.0 * alpha) / (2.0 * beta);
  for (i = 0; i + 1 < categories; ++i) cut[i] = incomplete_gamma(cut[i] * beta, alpha + 1, lnga1);
  output_rates[0] = cut[0] * mean * categories;
  output_rates[categories - 1] = (1 - cut[categories - 2]) * mean * categories;
  for (i = 1; i + 1 < categories; ++i) output_rates[i] = (cut[i] - cut[i - 1]) * mean * categories;
  free(cut);
  return BPPGPU_SUCCESS;
}
