/* galcamb.h -- C ABI of libgalcamb_b200.so: the B200 implementation of the scalar CMB transfer path of
 * the Galileon fork of CAMB (ClementLeloup/cosmomc_galileon).
 *
 * Convention = the reference's own iso_c_binding convention (camb/interface.f90:7-92, camb/galileon.cc):
 * trailing-underscore symbols, every scalar argument BY REFERENCE, arrays as plain pointers the caller owns
 * (Fortran: type(C_PTR), value + C_LOC(array(1))), integer status return (0 = ok, otherwise a CAMB error id of
 * camb/constants.f90:70-79), never exit().  All floating-point data are float64.  No CUDA / torch types appear.
 *
 * What each entry point replaces in the reference:
 *   galcamb_evolve_sources_   the OpenMP loop over source k-modes            camb/cmbmain.f90:201-206
 *                             (DoSourcek :662, CalcScalarSources :917, GaugeInterface_EvolveScal
 *                              camb/equations_galileon.f90:314, dverk camb/subroutines.f90:370, derivs/output)
 *   galcamb_los_integrate_    InitSourceInterpolation camb/cmbmain.f90:244 (:1267) and the OpenMP loop over
 *                             integration k-modes :263-267 (SourceToTransfers :483, InterpolateSources :1355,
 *                             DoFlatIntegration :1500)
 *   galcamb_scalar_cls_       CalcScalCls camb/cmbmain.f90:2167 + InterpolateCls :2450
 *   galcamb_set_bessels_      InitSpherBessels / GenerateBessels camb/bessels.f90:34-120
 *   galcamb_lens_cls_         lens_Cls -> CorrFuncFullSkyImpl camb/lensing.f90:78-528 (post-step, optional)
 *   galcamb_set_primordial_ + galcamb_transfers_to_powers_   CAMB_TransfersToPowers camb/camb.f90:87-102
 * Everything before (CAMBParams_Set, inithermo ...) and after (file output) stays in the Fortran host; the host
 * hands over the tables it has after InitVars through galcamb_model.  The Galileon background itself (arrays_ and
 * the rest of camb/galileon.cc) has a host-side replacement in the same library: include/galileon_abi.h.
 *
 * Threading: like CAMB itself (camb/cmbmain.f90:7-8) one context = one model batch at a time; call from one
 * host thread per GPU.  Device buffers are owned by the library, host buffers by the caller.
 */
#ifndef GALCAMB_H
#define GALCAMB_H
#ifdef __cplusplus
extern "C" {
#endif

#define GALCAMB_MAX_NU 5 /* = max_nu, camb/modules.f90:63 */
#define GALCAMB_MAX_NQ 5
#define GALCAMB_MAX_REGIONS 16
#define GALCAMB_MAX_TRANSFER_Z 16 /* transfer redshifts per call (reference: max_transfer_redshifts = 150, camb/modules.f90:64) */
#define GALCAMB_TRANSFER_MAX 13   /* rows of TransferData, camb/modules.f90:1876-1884 */

/* One region of a Ranges grid (camb/utils.F90:8-20). */
typedef struct galcamb_region {
  double Low, High, delta;
  int start_index; /* 1-based index of the first point of the region, as in the reference */
  int IsLog;
} galcamb_region;

/* The state the Fortran host holds after CAMBParams_Set + InitVars (camb/cmbmain.f90:722) for ONE model. */
typedef struct galcamb_model {
  /* camb/modules.f90:171-193 */
  double grhom, grhog, grhor, grhob, grhoc, grhov, grhornomass, grhok;
  double adotrad, tau0, taurst, taurend, tau_maxvis;
  double tight_tau, matter_verydom_tau;          /* ThermoData, camb/modules.f90:2717-2718 */
  double epsw;                                   /* GaugeInterface, camb/equations_galileon.f90:271 */
  double max_etak_scalar;                        /* camb/cmbmain.f90:752-753 */
  double AccuracyBoost, lAccuracyBoost;          /* camb/modules.f90:202-214 */
  int use_galileon, DoLateRadTruncation, HighAccuracyDefault, AccuratePolarization, AccurateReionization;
  int WantLateTime;                              /* camb/cmbmain.f90:137 ; selects 2 or 3 sources */
  int MassiveNuMethod;                           /* camb/modules.f90:56 */
  /* massive neutrinos: camb/modules.f90:186-189, :1570-1577 ; GaugeInterface :274 */
  int Nu_mass_eigenstates, Num_Nu_massive, nqmax;
  double grhormass[GALCAMB_MAX_NU], nu_masses[GALCAMB_MAX_NU];
  double nu_q[GALCAMB_MAX_NQ], nu_int_kernel[GALCAMB_MAX_NQ];
  double nu_tau_notmassless[GALCAMB_MAX_NQ][GALCAMB_MAX_NU]; /* [iq][eigenstate] */
  double nu_tau_nonrelativistic[GALCAMB_MAX_NU], nu_tau_massive[GALCAMB_MAX_NU];
  double dlnam;
  const double *nu_r1, *nu_p1, *nu_dr1, *nu_dp1, *nu_ddr1, *nu_ddp1; /* [2000] each, camb/modules.f90:1562-1566 */
  /* thermo tables on the log-tau grid, camb/modules.f90:2705-2714 */
  int nthermo;
  double tauminn, dlntau;
  const double *cs2, *dcs2, *dotmu, *ddotmu, *dddotmu; /* [nthermo] each */
  /* output time grid TimeSteps and the per-step visibility arrays, camb/modules.f90:2715, :3106-3146 */
  int nt;
  const double *tau, *dtau;                              /* TimeSteps%points, %dpoints */
  const double *vis, *dvis, *ddvis, *expmmu, *opac, *dopac;
  int n_tau_regions;
  galcamb_region tau_regions[GALCAMB_MAX_REGIONS];       /* TimeSteps%R(1:count), for Ranges_IndexOf */
  /* Galileon background owned by galileon.cc (camb/galileon.cc:72-92, :666-682) */
  double c2, c3, c4, c5, cG, h0, orad;
  int ngal;                                               /* the nb + 1 = 30 000 tabulated points (camb/galileon.cc:72) */
  const double *gal_a, *gal_h, *gal_x;                    /* ascending a; h = hbar(a); x = dpi/da */
  /* k grids */
  int nk;
  const double* k;                                        /* Evolve_q%points */
  int nq;
  const double *q, *dq;                                   /* ThisCT%q%points, %dpoints */
  /* primordial power, camb/power_tilt.f90:60-92 */
  double ScalarPowerAmp, an, n_run, n_runrun, k_0_scalar;
  int lmax;                                               /* CP%Max_l */
  /* matter transfer functions (CP%WantTransfer): output times tautf (camb/cmbmain.f90:783-793, ascending) and the
   * wavenumbers InitTransfer appends beyond the source grid, MT%q_trans(Evolve_q%npoints+1:) (camb/cmbmain.f90:597-623) */
  int WantTransfer, transfer_num_redshifts;
  double tautf[GALCAMB_MAX_TRANSFER_Z];
  int nk_transfer;
  const double* k_transfer;
  double H0;                                              /* CP%H0 (Transfer_kh = k/(H0/100)) */
  /* lensing amplitudes CosmoMC sets before every call (source/Calculator_CAMB.f90:130-131): ALens scales C_phiphi and,
   * by its root, C_phiT / C_phiE (camb/cmbmain.f90:122, :2254-2259); ALens_Fiducial > 0 replaces the lensing potential
   * by the scaled template (camb/lensing.f90:52, :252-257).  ALens = 0 in a zero-filled struct is read as 1. */
  double ALens, ALens_Fiducial;
} galcamb_model;

/* --- context --------------------------------------------------------------------------------------- */
int galcamb_init_(const int* device);
void galcamb_free_(void);
const char* galcamb_last_error_(void);
int galcamb_sizeof_model_(void); /* sizeof(galcamb_model) as compiled, for binding self-checks */
int galcamb_get_bessels_(int* nx, double* xs, double* ajl, double* ajlpr); /* x grid and tables [nl][nx] (tests) */

/* l samples (lSamp%l(1:l0)), Bessel x-range and accuracy: builds ajl/ajlpr on the device once and caches
 * them across models like camb/bessels.f90:40-41. */
int galcamb_set_bessels_(const int* ls, const int* nl, const int* kmaxfile, const double* AccuracyBoost);
/* fiducial template highL_CL_template(lmin:lmax_tmpl, 1:3) = TT,EE,TE (camb/modules.f90:1222-1245), laid out
 * [type][l-2]; pass lmax_tmpl = 0 for use_spline_template = F. */
int galcamb_set_template_(const double* tmpl, const int* lmax_tmpl);
/* 4th column of the same template file, the lensing potential [l(l+1)]^2 C_l^phiphi/2pi (camb/modules.f90:1236-1240),
 * [l-2]; only the lensing stage reads it (extrapolation beyond Max_l, camb/lensing.f90:237-253). */
int galcamb_set_lensing_template_(const double* pp, const int* lmax_tmpl);

/* --- batch of models ------------------------------------------------------------------------------- */
int galcamb_batch_begin_(const int* nmodels);
int galcamb_set_model_(const int* slot, const galcamb_model* m); /* slot = 0..nmodels-1; copies host -> device */

/* stage 1+2: Src(nk, S, nt) for every slot */
int galcamb_evolve_sources_(void);
/* stage 3: ddSrc and Delta_p_l_k(S, nl, nq) for every slot */
int galcamb_los_integrate_(void);
/* stage 4: iCl_scalar(nl, ntype) and Cl_scalar(lmin:lmax, ntype) for every slot */
int galcamb_scalar_cls_(void);

/* results (device -> host); layouts are the Fortran ones, column-major:
 *   Src(nk,S,nt)   Delta_p_l_k(S,nl,nq)   iCl(nl,ntype)   Cl(lmin:lmax, ntype), ntype = 3 (TT,EE,TE) or 6 */
int galcamb_get_sources_(const int* slot, double* Src);
int galcamb_put_sources_(const int* slot, const double* Src); /* feed stage 3 with host-computed sources */
int galcamb_get_transfers_(const int* slot, double* Delta_p_l_k);
int galcamb_get_cls_(const int* slot, double* iCl, double* Cl); /* either pointer may be NULL */
int galcamb_put_cls_(const int* slot, const double* Cl);         /* feed stage 5 with host-computed Cl_scalar */
int galcamb_get_status_(const int* slot, int* status);           /* per-slot CAMB error id */
/* Matter transfer functions of a slot after galcamb_evolve_sources_ (models with WantTransfer): MT%num_q_trans,
 * the number of redshifts, MT%q_trans, MT%TransferData as FP64 in Fortran order TransferData(13, num_q_trans, nz)
 * = C [nz][num_q_trans][13] (camb/modules.f90:1896-1915; rows 2..13 already divided by k^2 like outtransf,
 * camb/equations_galileon.f90:2093) and sigma_8 per redshift (Transfer_Get_sigmas, camb/modules.f90:2450).  Any
 * pointer may be NULL.  Replaces the transfer taps of camb/cmbmain.f90:1017-1068 and TransferOut (:1151-1201). */
int galcamb_get_transfer_(const int* slot, int* num_q_trans, int* num_redshifts, double* q_trans, double* TransferData,
                          double* sigma8);

/* stage 5 (post-step of the path): lens_Cls -> CorrFuncFullSky -> CorrFuncFullSkyImpl, camb/lensing.f90:78-528
 * (lensing_method = 1, AccurateBB = F, no tensors), on the device-resident Cl_scalar of every slot of the batch.
 * The models must carry the lensing potential (DoLensing: 3 sources) and share lmax and accuracy settings.
 * Returns 8 (error_norm_lensing, camb/constants.f90:80) if a slot fails the normalisation check (:230-235).
 *   Cl_lensed(lmin:lmax_lensed, 1:4) = TT, EE, BB, TE laid out [type][l-2] */
int galcamb_lens_cls_(void);
/* CAMB_TransfersToPowers, camb/camb.f90:87-102 -- what CosmoMC calls when only fast parameters changed
 * (source/Calculator_CAMB.f90:258): new InitPower values for a slot whose Delta_p_l_k is still resident, then
 * ClTransferToCl (and lens_Cls if *do_lensing) for the batch.  No ODE, no line-of-sight work. */
int galcamb_set_primordial_(const int* slot, const double* ScalarPowerAmp, const double* an, const double* n_run,
                            const double* n_runrun, const double* k_0_scalar);
/* the lensing amplitudes of a resident slot (fast parameters like the primordial ones): next galcamb_transfers_to_powers_ */
int galcamb_set_alens_(const int* slot, const double* ALens, const double* ALens_Fiducial);
int galcamb_transfers_to_powers_(const int* do_lensing);
int galcamb_get_lensed_cls_(const int* slot, int* lmax_lensed, double* Cl_lensed /* may be NULL */);

/* batch_begin + set_model for n models at once (tables packed on the host threads) */
int galcamb_batch_upload_(const galcamb_model* models, const int* nmodels);
/* one call for a whole batch: set models, run stages 1-4, return Cl(lmin:lmax,3) per model contiguous */
int galcamb_batch_cls_(const galcamb_model* models, const int* nmodels, double* Cl_out);
/* the same followed by stage 5: additionally returns Cl_lensed(lmin:lmax_lensed,4) per model contiguous (the caller
 * sizes it for lmax_lensed <= lmax - 100); Cl_out may be NULL.  Failed slots are NaN rows (status 4 or 8). */
int galcamb_batch_lensed_cls_(const galcamb_model* models, const int* nmodels, double* Cl_out, int* lmax_lensed,
                              double* Cl_lensed_out);

/* --- host pre-stage (csrc/prestage.cu) ------------------------------------------------------------------
 * The subset of CAMBparams (camb/modules.f90:98-167) and of the module-level accuracy knobs (:202-214) the scalar,
 * flat path reads; same names and meaning as the reference.  galcamb_default_params_ fills the defaults of
 * camb/inidriver.F90 / CAMB_SetDefParams (camb/camb.f90:214-269). */
typedef struct galcamb_params {
  double H0, omegab, omegac, omegav, omegan;     /* CP%H0, CP%omegab ... (fractions of the critical density) */
  double TCMB, YHe, Num_Nu_massless;
  int Num_Nu_massive, Nu_mass_eigenstates, share_delta_neff;
  double Nu_mass_degeneracies[GALCAMB_MAX_NU], Nu_mass_fractions[GALCAMB_MAX_NU];
  int Nu_mass_numbers[GALCAMB_MAX_NU];
  int use_galileon;                              /* camb/modules.f90:164-165 */
  double c2, c3, c4, cG;                         /* c5 / the highest non-zero coefficient follows from flatness */
  double ScalarPowerAmp, an, n_run, n_runrun, k_0_scalar; /* camb/power_tilt.f90:60-92 */
  int Reionization, use_optical_depth;           /* camb/reionization.f90:64-77 */
  double optical_depth, reion_redshift, delta_redshift, reion_fraction;
  double helium_redshift, helium_delta_redshift, helium_redshiftstart;
  double RECFAST_fudge, RECFAST_fudge_He;        /* camb/recfast.f90:257-282 (fudge already shifted for Hswitch) */
  int RECFAST_Heswitch, RECFAST_Hswitch;
  int Max_l;
  double Max_eta_k;
  int DoLensing, AccuratePolarization, AccurateReionization, MassiveNuMethod, DoLateRadTruncation, HighAccuracyDefault;
  double AccuracyBoost, lSampleBoost, lAccuracyBoost;
  int use_spline_template;
  /* CP%WantTransfer + TransferParams (camb/modules.f90:75-90): kmax in 1/Mpc, redshifts in descending order;
   * transfer_high_precision = T is not restated (error_unsupported_params) */
  int WantTransfer, transfer_high_precision, transfer_k_per_logint, transfer_num_redshifts;
  double transfer_kmax;
  double transfer_redshifts[GALCAMB_MAX_TRANSFER_Z];
  double ALens, ALens_Fiducial;                  /* camb/cmbmain.f90:122, camb/lensing.f90:52; defaults 1 and 0 */
} galcamb_params;

void galcamb_default_params_(galcamb_params* p);
int galcamb_sizeof_params_(void);
/* CAMBParams_Set (camb/modules.f90:260) + InitVars (camb/cmbmain.f90:722) + SetkValuesForSources (:797) +
 * SetkValuesForInt (:1281) + initlval (camb/modules.f90:851) for n parameter points, spread over *nthreads host threads
 * (0 = all cores).  The tables stay in the library until the next call / galcamb_prestage_free_; a point that fails
 * keeps its CAMB error id (galcamb_prestage_status_) and the call returns the first such id. */
int galcamb_prestage_(const galcamb_params* params, const int* n, const int* nthreads);
int galcamb_prestage_status_(const int* i, int* status);
const char* galcamb_prestage_error_(void);
/* a galcamb_model whose pointers look into the tables of point i (valid until the next pre-stage call) */
int galcamb_prestage_model_(const int* i, galcamb_model* m);
/* lSamp%l(1:l0) and the Bessel x range of point i: the arguments of galcamb_set_bessels_ (ls may be NULL to query nl) */
int galcamb_prestage_lsamples_(const int* i, int* nl, int* ls, int* kmaxfile);
void galcamb_prestage_free_(void);
/* pipelined form: _async_ starts the pre-stage of the NEXT batch on a second buffer and returns at once, _wait_ joins it
 * and makes that batch the current one; galcamb_prestaged_to_cls_ runs stages 1-4 on the current batch (Cl_out as in
 * galcamb_params_to_cls_, one row per point of the batch, NaN rows for rejected points). */
int galcamb_prestage_async_(const galcamb_params* params, const int* n, const int* nthreads);
int galcamb_prestage_wait_(void);
int galcamb_prestage_count_(void);
int galcamb_prestaged_to_cls_(double* Cl_out);
int galcamb_prestaged_upload_(void); /* current pre-staged batch -> device slots (then the stage calls apply) */
/* host threads the library may use (pre-stage, table packing); 0 = all cores.  One process per GPU: cores / ranks. */
void galcamb_set_host_threads_(const int* n);
int galcamb_get_host_threads_(void);
/* Where the Galileon background integrations of a batch run (arrays_, camb/galileon.cc:503-688): device >= 0 -- one lane
 * per cosmology in one launch per batch, next to the per-k kernel of the previous batch; -1 (default) -- on the host
 * threads.  Batches of one point always take the host path. */
int galcamb_set_prestage_device_(const int* device);
/* parameters -> C_l in one call: pre-stage on the host threads, stages 1-4 on the device.  Points the pre-stage or the
 * integration rejects come back as NaN rows; Cl_out as in galcamb_batch_cls_. */
int galcamb_params_to_cls_(const galcamb_params* params, const int* n, const int* nthreads, double* Cl_out);

/* --- several GPUs from one process -------------------------------------------------------------------------------
 * galcamb_init_ may be called once per device; a host thread works on the context of the device it named last
 * (galcamb_init_ / galcamb_select_device_), so OpenMP threads of a Fortran host can drive one GPU each.  The two calls
 * below do that themselves with one worker thread per device:
 *   parameter-batch sharding: contiguous blocks of the valid points per device, no exchange between devices;
 *   one spectrum over ngpu devices: source k-modes dealt round-robin in cost order, Src merged by an NCCL all-reduce
 *   (the one coupling of the k loop, camb/cmbmain.f90:1274), the integration k-modes sliced, iCl all-reduced.
 * ms (may be NULL) = {whole call, host pre-stage, slowest device's K1, slowest device's line-of-sight slice}. */
int galcamb_device_count_(void);
int galcamb_select_device_(const int* device);
int galcamb_multi_params_to_cls_(const int* ngpu, const galcamb_params* params, const int* n, const int* nthreads, double* Cl_out);
int galcamb_multi_single_cls_(const int* ngpu, const galcamb_params* params, double* Cl_out, double* ms);

/* measurement hooks (bench.py): last stage durations in ms (CUDA events on the library stream) and the
 * algorithmic work counters of SURVEY.md section 8(d). */
int galcamb_get_timings_(double* ms /*[8]: evolve, spline_k, los, cls, lensing, -, bessel, total of 0..3*/);
int galcamb_get_counters_(double* c /*[4]: n_derivs, sum_n, n_out, n_iter*/);
/* wall clock [ms] of the host-side phases of the last galcamb_batch_cls_ call: packing, H2D enqueue, upload as a whole,
   device stages 1-4, D2H of the C_l, collection of the pre-staged models, -, - */
int galcamb_get_host_timings_(double* ms /*[8]*/);
/* K1 executes fewer RHS evaluations than the reference algorithm counts in n_derivs: the y' of an output time
 * (camb/equations_galileon.f90:1297) is the k1 of the next step.  s[0] = evaluations executed, s[1] = outputs served so. */
int galcamb_get_evolve_stats_(double* s /*[2]*/);
int galcamb_sync_(void);
/* FP64 FMA micro-benchmark (roofline denominator): out[0] = device peak TFLOP/s, out[1], out[2] = one warp per SM with
 * 1 and 4 dependent chains (latency probes) */
int galcamb_fp64_peak_(double* out);

#ifdef __cplusplus
}
#endif
#endif
