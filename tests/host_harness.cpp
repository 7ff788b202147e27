// Host build of the device arithmetic (lfit_python_b200/csrc/roche_device.cuh compiled as plain
// C++) against the CPU oracle's independent C restatement.  Test tooling only: lets the device
// solver be checked on a machine without a GPU.  Prints "OK" or the first failure.
#include "../lfit_python_b200/csrc/roche_device.cuh"
extern "C" {
#include "../oracle/roche_core.h"
}
#include <stdio.h>
#include <stdlib.h>

static double urand() { return rand() / (RAND_MAX + 1.0); }

int main(int argc, char** argv)
{
    int N = argc > 1 ? atoi(argv[1]) : 50000;
    srand(12345);
    double worst = 0.0;
    for (double lq = -2.5; lq <= 1.0; lq += 0.02) {
        double q = pow(10.0, lq), x;
        lfb::Roche R;
        lfo_roche Ro;
        if (!lfb::roche_init(q, R) || lfo_xl1(q, &x) || lfo_roche_init(&Ro, q)) { printf("FAIL init q=%g\n", q); return 1; }
        worst = fmax(worst, fabs(R.xl1 - x));
        worst = fmax(worst, fabs(R.rin - 0.9 * lfo_lobe_polar_radius(&Ro)));
        if (q < 5) {
            double a = lfb::findphi90(R), b = lfo_findphi90(&Ro);
            worst = fmax(worst, fabs(a - b));
            for (double f = 0.05; f < 1; f += 0.1) {
                double s1, s2;
                bool ok1 = lfb::findi(R, f * a, a, s1);
                int ok2 = !lfo_findi(&Ro, f * b, b, &s2);
                if (ok1 != (ok2 != 0)) { printf("FAIL findi q=%g f=%g\n", q, f); return 1; }
                if (ok1) worst = fmax(worst, fabs(s1 - s2));
            }
        }
    }
    if (worst > 1e-13) { printf("FAIL scalars %g\n", worst); return 1; }
    int necl = 0;
    double md = 0.0;
    for (int t = 0; t < N; ++t) {
        double q = 0.03 + urand() * 0.97;
        if (urand() < 0.1) q = 1 + urand() * 2;
        lfb::Roche R;
        lfo_roche Ro;
        lfb::roche_init(q, R);
        lfo_roche_init(&Ro, q);
        // (a quarter of the elements at low inclinations, where lines of sight pass over the lobe's pole)
        double inc = (t % 4 == 0) ? 20 + urand() * 40 : 60 + urand() * 30, si = sin(inc * lfb::kDeg), ci = cos(inc * lfb::kDeg);
        lfb::Point T = {0, 0, 0, 0, 0};
        lfo_point To = {{0, 0, 0}, 0, 0};
        int kind = rand() % 3;
        if (kind == 1) {
            double r = urand() * 0.1 * R.xl1, al = urand() * 6.283;
            T.xi = To.xi = r * cos(al);
            T.eta = To.eta = r * sin(al);
        } else {
            double r = (0.02 + urand() * 0.6) * R.xl1, az = urand() * 6.283;
            T.x = To.p0[0] = r * cos(az);
            T.y = To.p0[1] = r * sin(az);
            if (kind == 2) T.z = To.p0[2] = (urand() - 0.5) * 0.05;
        }
        double a1, b1, a2, b2;
        // (elements in the orbital plane also through the solver's PLANAR form, which the disc and strip kernels use)
        int h1 = (kind == 0 && t % 2) ? lfb::ingress_egress<true>(R, si, ci, T, &a1, &b1) : lfb::ingress_egress(R, si, ci, T, &a1, &b1);
        int h2 = (t % 16 == 0) ? lfo_ingress_egress_robust(&Ro, si, ci, &To, &a2, &b2)
                               : lfo_ingress_egress_newton(&Ro, si, ci, &To, &a2, &b2);
        if (h1 != h2) { printf("FAIL eclipsed mismatch q=%g inc=%g\n", q, inc); return 1; }
        if (h1) { ++necl; md = fmax(md, fmax(fabs(a1 - a2), fabs(b1 - b2))); }
    }
    if (md > 1e-12 || necl < N / 3) { printf("FAIL ingress/egress maxdiff %g, eclipsed %d\n", md, necl); return 1; }
    // White-dwarf tiles solved from the centre's grazing lines of sight (the pipeline's path), over partial and
    // grazing eclipses of the white dwarf: short eclipses push the Newton iteration against the deepest line
    // of sight, where a halved step must not be taken for convergence.
    int ntile = 0;
    double mt = 0.0;
    for (int t = 0; t < N / 4; ++t) {
        double q = 0.05 + urand() * 0.9;
        lfb::Roche R;
        lfo_roche Ro;
        lfb::roche_init(q, R);
        lfo_roche_init(&Ro, q);
        double w90 = lfb::findphi90(R), s;
        double dphi = (t % 2 ? 0.02 + 0.3 * urand() * urand() : 0.2 + 0.8 * urand()) * w90;
        if (!lfb::findi(R, dphi, w90, s)) continue;
        double si = s, ci = sqrt(1.0 - s * s);
        lfb::Point T0 = {0, 0, 0, 0, 0};
        lfb::Roots hint;
        hint.lam[0] = NAN;
        double a0, b0;
        bool have = lfb::ingress_egress(R, si, ci, T0, &a0, &b0, nullptr, &hint) && hint.lam[0] == hint.lam[0];
        double rwd = (0.005 + 0.1 * urand()) * R.xl1;
        // every eighth walker: tiles that are only just eclipsed (eclipses down to 1e-4 of a cycle), found by
        // bisecting the offset along one direction on the sky between "eclipsed" and "not"
        double al_b = urand() * 6.283185307179586, r_b = -1.0;
        if (t % 8 == 1) {
            lfo_point Tb = {{0, 0, 0}, 0.12 * R.xl1 * cos(al_b), 0.12 * R.xl1 * sin(al_b)};
            double ab, bb;
            if (!lfo_ingress_egress_newton(&Ro, si, ci, &Tb, &ab, &bb)) {
                double lo = 0.0, hi = 0.12 * R.xl1;
                for (int it = 0; it < 30; ++it) {
                    double m = 0.5 * (lo + hi);
                    Tb.xi = m * cos(al_b);
                    Tb.eta = m * sin(al_b);
                    if (lfo_ingress_egress_newton(&Ro, si, ci, &Tb, &ab, &bb)) lo = m; else hi = m;
                }
                r_b = lo;
            }
        }
        for (int k = 0; k < 6; ++k) {
            double r = rwd * sqrt(urand()), al = urand() * 6.283185307179586;
            if (r_b > 0.0) { r = r_b * (1.0 - pow(10.0, -1.0 - k)); al = al_b; }
            lfb::Point T = {0, 0, 0, r * cos(al), r * sin(al)};
            lfo_point To = {{0, 0, 0}, T.xi, T.eta};
            double a1, b1, a2, b2;
            int h1 = lfb::ingress_egress(R, si, ci, T, &a1, &b1, have ? &hint : nullptr);
            int h2 = lfo_ingress_egress_robust(&Ro, si, ci, &To, &a2, &b2);
            if (h1 != h2) { printf("FAIL tile eclipsed mismatch q=%.17g dphi=%.17g xi=%.17g eta=%.17g\n", q, dphi, T.xi, T.eta); return 1; }
            if (h1) {
                ++ntile;
                double d = fmax(fabs(a1 - a2), fabs(b1 - b2));
                // (near the double root an error eps of the potential moves a boundary by ~ eps / width)
                if (d > 1e-11 + 4e-16 / fabs(b2 - a2)) { printf("FAIL tile q=%.17g dphi=%.17g xi=%.17g eta=%.17g: %.12f %.12f vs %.12f %.12f\n", q, dphi, T.xi, T.eta, a1, b1, a2, b2); return 1; }
                if (fabs(b2 - a2) > 1e-4) mt = fmax(mt, d);
            }
        }
    }
    if (ntile < N / 4) { printf("FAIL too few eclipsed tiles %d\n", ntile); return 1; }
    double mb = 0.0;
    for (int t = 0; t < 300; ++t) {
        double q = 0.03 + urand() * 0.97;
        lfb::Roche R;
        lfo_roche Ro;
        lfb::roche_init(q, R);
        lfo_roche_init(&Ro, q);
        double rad = (0.05 + 0.7 * urand()) * R.xl1, o1[4], o2[4];
        bool ok1 = lfb::bspot(R, rad, o1);
        int ok2 = !lfo_bspot(&Ro, rad, o2);
        if (ok1 != (ok2 != 0)) { printf("FAIL bspot mismatch q=%g rad=%g\n", q, rad); return 1; }
        if (ok1) for (int j = 0; j < 4; ++j) mb = fmax(mb, fabs(o1[j] - o2[j]));
    }
    if (mb > 1e-10) { printf("FAIL bspot maxdiff %g\n", mb); return 1; }
    double mr = 0.0;
    for (int t = 0; t < 5000; ++t) {
        double q = 0.03 + urand() * 2;
        lfb::Roche R;
        lfo_roche Ro;
        lfb::roche_init(q, R);
        lfo_roche_init(&Ro, q);
        double th = 0.04 + urand() * (lfb::kPi - 0.04), ph = urand() * 6.283;
        double dx = -cos(th), dy = sin(th) * cos(ph), dz = sin(th) * sin(ph), g[3];
        double r = lfb::donor_radius(R, dx, dy, dz, g);
        double lo = 0.02 * Ro.rs, hi = Ro.rs;
        for (int it = 0; it < 100; ++it) {
            double m = 0.5 * (lo + hi);
            if (lfo_pot(&Ro, 1 + m * dx, m * dy, m * dz) < Ro.phic) lo = m; else hi = m;
        }
        mr = fmax(mr, fabs(r - 0.5 * (lo + hi)));
    }
    if (mr > 1e-11) { printf("FAIL donor radius maxdiff %g\n", mr); return 1; }
    printf("OK scalars %.1e ie %.1e (%d eclipsed) tiles %.1e (%d) bspot %.1e donor %.1e\n", worst, md, necl, mt, ntile, mb, mr);
    return 0;
}
