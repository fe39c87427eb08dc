#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static void sc(double x, double* s, double* c) {
  const double k = rint(x * 0.63661977236758134308);
  double r = fma(-k, 1.57079632673412561417e+00, x);
  r = fma(-k, 6.07710050650619224932e-11, r);
  const double z = r * r;
  const double ps = fma(z, fma(z, fma(z, fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08), 2.75573137070700676789e-06), -1.98412698298579493134e-04), 8.33333333332248946124e-03);
  const double sr = fma(z * r, fma(z, ps, -1.66666666666666324348e-01), r);
  const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09), -2.75573143513906633035e-07), 2.48015872894767294178e-05), -1.38888888888741095749e-03), 4.16666666666666019037e-02);
  const double cr = 1.0 - fma(0.5, z, -(z * z * pc));
  switch ((int)k & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}
int main() {
  const float factorPI = (float)(3.14159265358979323846 / 180.0);
  long bad = 0, n = 0; double maxe = 0;
  for (uint32_t bits = 0; bits <= 0x43b40000u; bits += 7) {  // floats 0 .. 360
    float ang; memcpy(&ang, &bits, 4);
    if (!(ang < 360.f)) break;
    float th = ang * factorPI;
    double s, c; sc((double)th, &s, &c);
    double s0 = sin((double)th), c0 = cos((double)th);
    if ((float)s != (float)s0 || (float)c != (float)c0) { if (bad < 5) printf("mismatch ang %.9g: %.17g %.17g | %.17g %.17g\n", ang, s, s0, c, c0); bad++; }
    double e = fabs(s - s0) > fabs(c - c0) ? fabs(s - s0) : fabs(c - c0); if (e > maxe) maxe = e;
    n++;
  }
  printf("n=%ld bad=%ld maxerr=%.3g\n", n, bad, maxe);
}
