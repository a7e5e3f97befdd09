// Device-side view of one cosmological model (everything the per-k perturbation kernel and the
// line-of-sight kernels read).  Plain data; built by capi.cu from the caller's galcamb_model.
//
// Layout in HBM (one block per model, all float64 unless noted):
//   th   [nthermo][5]   = {cs2, dcs2, dotmu, ddotmu, dddotmu} rows of the log-tau thermo table (AoS so one
//                          Hermite lookup touches 2 adjacent 40-byte rows instead of 10 separate arrays;
//                          reference tables: camb/modules.f90:2705-2713)
//   ts   [nt][8]        = {tau, dtau, vis, dvis, ddvis, expmmu, opac, dopac} per output time
//                          (camb/modules.f90:2715, TimeSteps%points/dpoints)
//   gal  [ngal][5]      = {a, hbar, c_h, x, c_x} rows of the Galileon background spline
//                          (camb/galileon.cc:666-682; c_* = natural-spline second derivative / 2)
//   dotmu_pmin [nthermo] = running minimum of dotmu (turns the linear scan of Thermo_OpacityToTime,
//                          camb/modules.f90:2771-2783, into a bisection with the same answer)
//   k    [nk]           = source wavenumbers (Evolve_q%points)
#pragma once
#include <cstddef>

#ifndef GC_MAX_NU
#define GC_MAX_NU 5      // = the reference's max_nu (camb/modules.f90:63)
#endif
#define GC_MAX_NQ 5      // momentum nodes (nqmax <= 5 for AccuracyBoost <= 3; camb/modules.f90:1624-1627)
#define GC_NRHOPN 2000   // camb/modules.f90:1554
#define GC_MAX_REGIONS 16
#define GC_MAX_TF 16      // transfer redshifts per call (GALCAMB_MAX_TRANSFER_Z; the reference allows 150, CosmoMC uses 1-4)
#define GC_TRANSFER_MAX 13  // rows of TransferData, camb/modules.f90:1876-1884

struct GcRegion {  // camb/utils.F90:8-20 (piecewise linear/log grid)
  double Low, High, delta;
  int start_index, IsLog;
};

struct GcRegions {
  int count, npoints;
  double Lowest, Highest;
  GcRegion R[GC_MAX_REGIONS];
};

struct DevModel {
  // ---- hot block: everything derivs() reads, packed into the first ~3 cache lines so that a warp whose 32
  // lanes belong to 32 different models keeps its constants L1 resident (12 warps x 32 models x ~350 B)
  double grhom, grhog, grhob, grhoc, grhov, grhornomass, grhok;   // camb/modules.f90:171-193
  double c2, c3, c4, c5, cG, h0, orad;                                   // camb/galileon.cc:80-92
  double tauminn, dlntau, dlnam;
  double inv_tauminn, inv_dlntau, inv_dlnam;  // reciprocals used by the table-index computations
  double inv_nu_q[GC_MAX_NQ];
  double gal_lnq_inv, gal_a0;  // index guess into the background table: i ~ log(a/a0)*gal_lnq_inv
  double grhormass[GC_MAX_NU], nu_masses[GC_MAX_NU];
  double nu_int_kernel[GC_MAX_NQ];
  const double* th;
  const double* gal;
  const double* nu_tab;  // [GC_NRHOPN][6] = {r1, dr1, p1, dp1, ddr1, ddp1} rows (shared by all models)
  int use_galileon, n_eig, Num_Nu_massive, nqmax, nthermo, ngal;
  // ---- everything else
  double grhor, nu_q[GC_MAX_NQ];
  double adotrad, tau0, taurst, taurend, tau_maxvis, tight_tau, matter_verydom_tau, epsw;
  double max_etak_scalar, AccuracyBoost, lAccuracyBoost, tol1;
  int DoLateRadTruncation, HighAccuracyDefault, AccuratePolarization, AccurateReionization;
  int WantLateTime, SourceNum, NuMethod;
  double nu_tau_notmassless[GC_MAX_NQ][GC_MAX_NU], nu_tau_nonrelativistic[GC_MAX_NU], nu_tau_massive[GC_MAX_NU];
  const double* dotmu_pmin;
  // time steps
  int nt;
  const double* ts;
  // source k grid and output
  int nk;
  const double* k;
  double* Src;    // [nt][S][nk]   (== Fortran Src(nk,S,nt))
  double* ddSrc;  // same shape, second derivatives along k
  // line-of-sight stage
  GcRegions TimeSteps;
  int nq;
  const double* q;    // [nq]
  const double* dq;   // [nq]
  double* Delta;      // [nq][nl][S]  (== Fortran Delta_p_l_k(S,nl,nq))
  // primordial power + C_l
  double ScalarPowerAmp, an, n_run, n_runrun, k_0_scalar;
  double* iCl;        // [ntype][nl]
  double* Cl;         // [ntype][lmax-1]
  // matter transfer functions (camb/cmbmain.f90:508-632, :1017-1068, :1151-1201)
  int WantTransfer, ntf, nk_tr;  // nk_tr = wavenumbers beyond the source grid (evolved through the transfer times only)
  double tautf[GC_MAX_TF];       // conformal times of the transfer redshifts, ascending
  double h0_100;                 // CP%h0/100 (Transfer_kh = k/h)
  const double* k_tr;            // [nk_tr]
  double* Transfer;              // [ntf][nk + nk_tr][13] FP64  (== Fortran TransferData(13, num_q_trans, nz))
  // lensing amplitude knobs CosmoMC sets before every call (source/Calculator_CAMB.f90:130-131)
  double ALens;      // camb/cmbmain.f90:122, :2254-2259
  double ALens_fid;  // ALens_Fiducial, camb/lensing.f90:52, :252-257 (0 = off)
};
