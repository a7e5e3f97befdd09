/* galileon_abi.h -- the reference's own C ABI for the Galileon background and perturbation terms, exported by
 * libgalcamb_b200.so as a replacement of camb/galileon.cc (no GSL, no call-backs into Fortran).
 *
 * Convention = camb/interface.f90:7-92: trailing-underscore names, every scalar by reference (GetdHdX_ returns dh, dx
 * through C++ references exactly like camb/galileon.cc:746), arrays as plain pointers the caller owns, integer status
 * 0 = OK (non-zero -> GlobalError(..., error_background_evolution), camb/equations_galileon.f90:97-100).
 * hcamb = conformal Hubble rate a*hbar in units of H0, xcamb = a dpi/da, as in the Fortran call sites
 * (camb/equations_galileon.f90:131-156, :2285-2310).  State is process-global like the reference's file-scope
 * tables (camb/galileon.cc:72-92): arrays_ rebuilds it, freegal_ drops it; the per-point functions only read it and
 * may be called concurrently (the OMP k loop, camb/cmbmain.f90:201).
 */
#ifndef GALILEON_ABI_H
#define GALILEON_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

/* Massive-neutrino tables of the host (camb/modules.f90:1595-1719: r1, p1, dr1, dp1, ddr1, ddp1 on the nrhopn = 2000
 * point log(am) grid with spacing dlnam); replaces the call-backs MassiveNu::Nu_rho / Nu_background / Nu_dp that
 * camb/galileon.cc:96-99 makes into Fortran.  Call once before arrays_ when there are massive eigenstates. */
int galileon_set_nu_tables_(const double* r1, const double* p1, const double* dr1, const double* dp1, const double* ddr1,
                            const double* ddp1, const int* n, const double* dlnam);

/* camb/galileon.cc:503 -- integrates hbar(a), x(a) from a = 1 down to 1e-10 on 30 000 log-spaced nodes, runs the
 * ghost/Laplace stability checks (:130-245) and the Friedmann closure test (:420) at every node, closes the highest
 * non-zero c_i by flatness (:520-575).  H0in = H0/c in Mpc^-1.  Status: 0 OK, 4 closure, 7 unstable, 8 NaN, 5 bad input */
int arrays_(double* omegar, double* omegam, double* H0in, double* c2in, double* c3in, double* c4in, double* cGin,
            double* grhormass, double* nu_masses, int* nu_mass_eigenstates);
void freegal_(void);                                             /* camb/galileon.cc:1278 */
double GetX_(double* point);                                     /* :691  x(a) = a dpi/da      (NaN outside [0, 1.1]) */
double GetH_(double* point);                                     /* :719  hbar(a) = H/H0 */
#ifdef __cplusplus
void GetdHdX_(double* point, double* hcamb, double* xcamb, double& dh, double& dx, double* grhormass, double* nu_masses,
              int* nu_mass_eigenstates);                         /* :746  d/dlna of hbar-conformal and x */
#endif
double ct_(double* point, double* grhormass, double* nu_masses, int* nu_mass_eigenstates); /* :789 tensor speed */
double grhogal_(double* point, double* hcamb, double* xcamb);                              /* :809 */
double gpresgal_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb); /* :831 */
double Chigal_(double* point, double* hcamb, double* xcamb, double* dgrho, double* eta, double* dphi, double* dphiprime,
               double* k);                                                                 /* :852 */
double qgal_(double* point, double* hcamb, double* xcamb, double* dgq, double* eta, double* dphi, double* dphiprime,
             double* k);                                                                   /* :897 */
double Pigal_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb, double* dgrho, double* dgq,
              double* dgpi, double* eta, double* dphi, double* k);                         /* :938 */
double dphisecond_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb, double* dgrho, double* dgq,
                   double* eta, double* dphi, double* dphiprime, double* k, double* deltafprime); /* :987 */
double pigalprime_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb, double* dgrho, double* dgq,
                   double* dgpi, double* pidot, double* eta, double* dphi, double* dphiprime, double* k, double* grho,
                   double* gpres, double* grhormass, double* nu_masses, int* nu_mass_eigenstates); /* :1084 */

/* camb/theta.cc:296 -- theta = r_s(a*)/D_A(a*) for CosmoMC's theta <-> H0 solve (source/CosmologyParameterizations.f90:
 * 263-271).  h0 in km/s/Mpc.  Runs its own background integration (arrays_' tables are untouched); -1 = rejected. */
double theta_(double* omegam, double* orad, double* ombh2, double* h0, double* astar, double* c2in, double* c3in, double* c4in,
              double* cGin, double* grhormass, double* nu_masses, int* nu_mass_eigenstates);

/* Not in the reference: the tabulated background (ascending a, 30 000 nodes) and the closed coefficients, i.e. what
 * galcamb_model.gal_a/gal_h/gal_x and c2..c5,cG are filled from. */
int galgettable_(const double** a, const double** h, const double** x, int* n);
int galgetcoefficients_(double* c2, double* c3, double* c4, double* c5, double* cG);

#ifdef __cplusplus
}
#endif
#endif
