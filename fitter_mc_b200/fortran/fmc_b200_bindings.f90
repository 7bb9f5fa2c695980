MODULE fmc_b200_bindings
  ! Fortran 2003 ISO_C_BINDING interface to libfmc_b200.so (include/fmc_b200.h).
  !
  ! This is the host-side layer BASELINE.json's north star asks for: fitter_mc stays a
  ! Fortran 2003 program and calls the CUDA library through bind(C).  It is shipped as
  ! source: no Fortran compiler exists in the build image (SURVEY.md section 0), so it is
  ! compiled where gfortran is available:
  !     gfortran -c fmc_b200_bindings.f90
  !     gfortran ... bootstrap.o fmc_b200_bindings.o -L<repo>/fitter_mc_b200 -lfmc_b200
  ! See INTEGRATION.md for the change to bootstrap.f90 (lines 297-330 become one call).
  USE, INTRINSIC :: iso_c_binding
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: fmc_bootstrap_run, fmc_bootstrap_run_binned, fmc_fit_once, fmc_generate_offsets, &
    &fmc_strerror_f, fmc_model_count, fmc_set_jacobian_mode, fmc_histogram_quantiles, &
    &fmc_set_schedule, fmc_last_combine_method, fmc_default_seed, &
    &fmc_schedule_refill, fmc_schedule_rounds, fmc_schedule_auto

  INTEGER(c_long_long), PARAMETER :: fmc_default_seed = 31415_c_long_long ! utils.f90:7
  ! fmc_set_schedule (include/fmc_b200.h): how the iterations of the fits are scheduled on the GPU
  INTEGER(c_int), PARAMETER :: fmc_schedule_refill = 0_c_int, fmc_schedule_rounds = 1_c_int, &
    &fmc_schedule_auto = 2_c_int

  INTERFACE

    ! int fmc_model_count(void)
    FUNCTION fmc_model_count() BIND(C, name='fmc_model_count') RESULT(n)
      IMPORT :: c_int
      INTEGER(c_int) :: n
    END FUNCTION fmc_model_count

    ! Replaces the body of the Monte Carlo loop, bootstrap.f90:297-330.
    FUNCTION fmc_bootstrap_run(model, ndata, x, y, err_y, params_start, nresample, seed, &
      &first_counter, ngpus, z_inject, sum_p, sum_p2, sum_prop, sum_prop2, params_out, &
      &props_out, rc_count, totals) BIND(C, name='fmc_bootstrap_run') RESULT(rc)
      IMPORT :: c_int, c_double, c_long_long, c_ptr
      INTEGER(c_int), VALUE :: model, ndata, ngpus
      REAL(c_double), INTENT(in) :: x(*), y(*), err_y(*), params_start(*)
      INTEGER(c_long_long), VALUE :: nresample, seed, first_counter
      TYPE(c_ptr), VALUE :: z_inject              ! c_null_ptr: use the generator
      REAL(c_double), INTENT(out) :: sum_p(*), sum_p2(*), sum_prop(*), sum_prop2(*)
      TYPE(c_ptr), VALUE :: params_out, props_out ! c_null_ptr or c_loc(array)
      TYPE(c_ptr), VALUE :: rc_count, totals      ! c_null_ptr or c_loc(INTEGER(c_long_long) array)
      INTEGER(c_int) :: rc
    END FUNCTION fmc_bootstrap_run

    ! The same loop with make_histogram=.TRUE. for large nrand_points: instead of one
    ! WRITE(8+j,*)props(j) per resample (bootstrap.f90:319-325) every parameter and property is
    ! binned on the GPU; counts is INTEGER(c_long_long) (nbins+2, nparam+nprop), first and last
    ! row of a column = resamples below bin_lo / at or above bin_hi.
    FUNCTION fmc_bootstrap_run_binned(model, ndata, x, y, err_y, params_start, nresample, seed, &
      &first_counter, ngpus, sum_p, sum_p2, sum_prop, sum_prop2, rc_count, totals, nbins, &
      &bin_lo, bin_hi, counts) BIND(C, name='fmc_bootstrap_run_binned') RESULT(rc)
      IMPORT :: c_int, c_double, c_long_long, c_ptr
      INTEGER(c_int), VALUE :: model, ndata, ngpus, nbins
      REAL(c_double), INTENT(in) :: x(*), y(*), err_y(*), params_start(*), bin_lo(*), bin_hi(*)
      INTEGER(c_long_long), VALUE :: nresample, seed, first_counter
      REAL(c_double), INTENT(out) :: sum_p(*), sum_p2(*), sum_prop(*), sum_prop2(*)
      TYPE(c_ptr), VALUE :: rc_count, totals
      INTEGER(c_long_long), INTENT(out) :: counts(*)
      INTEGER(c_int) :: rc
    END FUNCTION fmc_bootstrap_run_binned

    ! README:35-42 "replace the call to nl2sno with a call to nl2sol": mode 1 fits with exact
    ! Jacobian rows (the plug-in's model_grad, or dual numbers through its generic model_y).
    ! ctx = c_null_ptr sets the mode of fmc_bootstrap_run / fmc_fit_once.
    FUNCTION fmc_set_jacobian_mode(ctx, mode) BIND(C, name='fmc_set_jacobian_mode') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: ctx
      INTEGER(c_int), VALUE :: mode
      INTEGER(c_int) :: rc
    END FUNCTION fmc_set_jacobian_mode

    ! No counterpart in the reference: one persistent kernel with refill, or solver steps and passes as
    ! separate kernels in rounds; fmc_schedule_auto (the default) chooses by measurement.  Results agree
    ! to the parity floor.  ctx = c_null_ptr sets it for fmc_bootstrap_run / fmc_fit_once.
    FUNCTION fmc_set_schedule(ctx, schedule) BIND(C, name='fmc_set_schedule') RESULT(rc)
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: ctx
      INTEGER(c_int), VALUE :: schedule
      INTEGER(c_int) :: rc
    END FUNCTION fmc_set_schedule

    ! How the last fmc_bootstrap_run with ngpus > 1 combined the GPUs' moment blocks: 1 = one
    ! ncclAllReduce over NVLink (the REDUCTION(+) of bootstrap.f90:310-311), 0 = one GPU, 2 = summed on the host.
    FUNCTION fmc_last_combine_method() BIND(C, name='fmc_last_combine_method') RESULT(method)
      IMPORT :: c_int
      INTEGER(c_int) :: method
    END FUNCTION fmc_last_combine_method

    ! Quantiles of one quantity from its column of counts.
    FUNCTION fmc_histogram_quantiles(counts, nbins, lo, hi, probs, nprobs, out) &
      &BIND(C, name='fmc_histogram_quantiles') RESULT(outside)
      IMPORT :: c_int, c_double, c_long_long
      INTEGER(c_long_long), INTENT(in) :: counts(*)
      INTEGER(c_int), VALUE :: nbins, nprobs
      REAL(c_double), VALUE :: lo, hi
      REAL(c_double), INTENT(in) :: probs(*)
      REAL(c_double), INTENT(out) :: out(*)
      INTEGER(c_long_long) :: outside
    END FUNCTION fmc_histogram_quantiles

    ! offset_data(.FALSE.) ; fit_data(params)  -- bootstrap.f90:288-291
    FUNCTION fmc_fit_once(model, ndata, x, y, err_y, params, rc2) &
      &BIND(C, name='fmc_fit_once') RESULT(rc)
      IMPORT :: c_int, c_double
      INTEGER(c_int), VALUE :: model, ndata
      REAL(c_double), INTENT(in) :: x(*), y(*), err_y(*)
      REAL(c_double), INTENT(inout) :: params(*)
      INTEGER(c_int), INTENT(out) :: rc2(2)
      INTEGER(c_int) :: rc
    END FUNCTION fmc_fit_once

    ! The Gaussian offsets of resamples [first_counter, first_counter+nresample)
    FUNCTION fmc_generate_offsets(seed, first_counter, nresample, ndata, z_out) &
      &BIND(C, name='fmc_generate_offsets') RESULT(rc)
      IMPORT :: c_int, c_double, c_long_long
      INTEGER(c_long_long), VALUE :: seed, first_counter, nresample
      INTEGER(c_int), VALUE :: ndata
      REAL(c_double), INTENT(out) :: z_out(*)
      INTEGER(c_int) :: rc
    END FUNCTION fmc_generate_offsets

    FUNCTION fmc_strerror(code) BIND(C, name='fmc_strerror') RESULT(msg)
      IMPORT :: c_int, c_ptr
      INTEGER(c_int), VALUE :: code
      TYPE(c_ptr) :: msg
    END FUNCTION fmc_strerror

  END INTERFACE

CONTAINS

  FUNCTION fmc_strerror_f(code) RESULT(msg)
    ! Error text as a Fortran string, for  CALL errstop('MONTE_CARLO_FIT_DATA', ...)
    INTEGER(c_int), INTENT(in) :: code
    CHARACTER(:), ALLOCATABLE :: msg
    CHARACTER(kind=c_char), POINTER :: p(:)
    INTEGER :: n
    CALL c_f_pointer(fmc_strerror(code), p, [256])
    n = 0
    DO WHILE (n < 256)
      IF (p(n+1) == c_null_char) EXIT
      n = n + 1
    ENDDO
    ALLOCATE(CHARACTER(n) :: msg)
    msg = TRANSFER(p(1:n), msg)
  END FUNCTION fmc_strerror_f

END MODULE fmc_b200_bindings
