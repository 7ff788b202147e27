// Host build of the device GP likelihood (lfit_python_b200/csrc/gp_device.cuh as plain C++),
// exported with a C entry point so that tests can compare it with the oracle's dense Cholesky
// on a machine without a GPU.  Test tooling only.
#include "../lfit_python_b200/csrc/gp_device.cuh"

extern "C" double host_gp_loglike(int n, const double* x, const double* ye, const double* r, double a_in, double a_out,
                                  double tau, int n_gaps, const double* gaps)
{
    lfb::GpPars G;
    G.a_in = a_in;
    G.a_out = a_out;
    G.tau = tau;
    G.n_gaps = n_gaps;
    for (int k = 0; k < n_gaps; ++k) {
        G.gap[k][0] = gaps[2 * k];
        G.gap[k][1] = gaps[2 * k + 1];
    }
    return lfb::gp_loglike(
        n, [&](int k) { return x[k]; }, [&](int k) { return ye[k] * ye[k]; }, [&](int k) { return r[k]; }, G);
}

extern "C" double host_gp_loglike_two_sided(int n, const double* x, const double* ye, const double* r, double a_in,
                                            double a_out, double tau, int n_gaps, const double* gaps)
{
    lfb::GpPars G;
    G.a_in = a_in;
    G.a_out = a_out;
    G.tau = tau;
    G.n_gaps = n_gaps;
    for (int k = 0; k < n_gaps; ++k) {
        G.gap[k][0] = gaps[2 * k];
        G.gap[k][1] = gaps[2 * k + 1];
    }
    return lfb::gp_loglike_two_sided(
        n, [&](int k) { return x[k]; }, [&](int k) { return ye[k] * ye[k]; }, [&](int k) { return r[k]; }, G);
}

extern "C" int host_gp_changepoints(double x_min, double x_max, double dist_cp, double phi0, double* gaps)
{
    lfb::GpPars G;
    lfb::gp_changepoints(x_min, x_max, dist_cp, phi0, G);
    for (int k = 0; k < G.n_gaps; ++k) {
        gaps[2 * k] = G.gap[k][0];
        gaps[2 * k + 1] = G.gap[k][1];
    }
    return G.n_gaps;
}
