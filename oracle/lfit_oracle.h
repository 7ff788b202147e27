/*
 * oracle/lfit_oracle.h -- CPU FP64 oracle for the LFIT CV eclipse model.
 *
 * TEST INFRASTRUCTURE ONLY (see roche_core.h).  PARITY UNPINNED: `lfit` and
 * `trm.roche` are not under /root/reference; this restates the model that the
 * reference's call sites describe (CVModel.py:128-178, README.md:24-43,
 * testCV.py:17-65) with the discretisation written down in DESIGN.md.
 */
#ifndef LFIT_ORACLE_H
#define LFIT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* CV parameter order = SimpleEclipse.cv_parnames / ComplexEclipse.cv_parnames
 * (/root/reference/CVModel.py:326-333, 383-390) */
enum {
    LFO_WDFLUX = 0, LFO_DFLUX, LFO_SFLUX, LFO_RSFLUX, LFO_Q, LFO_DPHI, LFO_RDISC, LFO_ULIMB,
    LFO_RWD, LFO_SCALE, LFO_AZ, LFO_FIS, LFO_DEXP, LFO_PHI0, LFO_EXP1, LFO_EXP2, LFO_TILT,
    LFO_YAW, LFO_NPAR
};

/* flags for lfo_calc_flux */
enum {
    LFO_FLAG_INCL = 1,     /* slot LFO_DPHI holds the inclination in degrees instead of dphi */
    LFO_FLAG_SKIP_WD = 2, LFO_FLAG_SKIP_DISC = 4, LFO_FLAG_SKIP_BS = 8, LFO_FLAG_SKIP_DONOR = 16
};

enum { LFO_SOLVER_ROBUST = 0, LFO_SOLVER_NEWTON = 1 };

typedef struct {
    int n_wd_rings;   /* white dwarf: 4*n^2 equal-area sky tiles (default 10 -> 400) */
    int n_disc_r;     /* disc: radial rings (default 25) */
    int n_disc_th;    /* disc: azimuthal sectors, even (default 40 -> 1000 tiles, testCV.py:31) */
    int n_bs;         /* bright-spot strip elements (default 200) */
    int n_donor_th;   /* donor: rings in polar angle about the line of centres (default 18 -> 412 tiles, testCV.py:43) */
    int n_quad;       /* exposure quadrature points, odd (default 3: Simpson) */
    double donor_ulimb;  /* donor linear limb darkening (default 0.8) */
    double donor_gdexp;  /* donor gravity-darkening exponent on |grad Phi| (default 0.32 = 4*0.08) */
    int solver;       /* LFO_SOLVER_* */
} lfo_config;

void lfo_default_config(lfo_config *cfg);

/* trm.roche equivalents; return 0 on success, non-zero where the reference raises */
int lfo_roche_xl1(double q, double *out);
int lfo_roche_findphi(double q, double incl_deg, double *out);
int lfo_roche_findi(double q, double dphi, double *incl_deg);
int lfo_roche_bspot(double q, double rad, double out[4]);
int lfo_roche_ingress_egress(double q, double incl_deg, const double p0[3], double xi, double eta,
                             int solver, double *ph_in, double *ph_out);

/* lfit.CV.calcFlux: returns 0 and fills out_total[n_ph] (+ optional scaled component curves
 * ywd/yd/ys/yrs); non-zero (outputs NaN) where the parameters admit no model */
int lfo_calc_flux(const lfo_config *cfg, const double *pars, int npars, int flags, int n_ph,
                  const double *phase, const double *width, double *out_total, double *out_wd,
                  double *out_d, double *out_s, double *out_rs);

/* SimpleEclipse.chisq (CVModel.py:157-178): +inf for an invalid model */
double lfo_chisq(const lfo_config *cfg, const double *pars, int npars, int n_ph, const double *phase,
                 const double *width, const double *y, const double *ye);

/* Prior.ln_prob (model.py:83-113) */
enum { LFO_PRIOR_GAUSS = 0, LFO_PRIOR_GAUSSPOS, LFO_PRIOR_UNIFORM, LFO_PRIOR_LOGUNIFORM, LFO_PRIOR_MODJEFF };
double lfo_prior_ln_prob(int type, double p1, double p2, double norm, double val);

/* Flattened model tree (what Node.ln_prob walks, model.py:476-498) */
typedef struct {
    int ndim;              /* length of a walker's parameter vector */
    int n_ecl;             /* eclipses (leaves) */
    int npars;             /* 14 (simple BS) or 18 (complex) */
    const int *gather;     /* [n_ecl*18]: >=0 column of theta; <0: -(k+1) -> consts[k] */
    const double *consts;
    int n_prior;           /* every Param of the tree, variable or not */
    const int *prior_src;  /* same encoding as gather */
    const int *prior_type;
    const double *prior_p1, *prior_p2, *prior_norm;
    const int *prior_isvar;
    const long long *lc_off; /* [n_ecl+1] offsets into the concatenated lightcurves */
    const double *lc_phase, *lc_width, *lc_y, *lc_ye;
} lfo_layout;

/* what = 0 ln_prior, 1 ln_like, 2 ln_prob; theta row-major [n][ndim]; out[n];
 * chisq_out (optional) [n][n_ecl].  nthreads <= 0: all OpenMP threads. */
int lfo_log_prob(const lfo_config *cfg, const lfo_layout *L, int what, long long n, const double *theta,
                 double *out, double *chisq_out, int nthreads);

int lfo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
