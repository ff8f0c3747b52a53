/* oracle/shim/libint2/util/generated/libint2_params.h -- TEST INFRASTRUCTURE (oracle).
 *
 * Stand-in for the generator-emitted libint2_params.h (emitted by
 * /root/reference/src/bin/libint/build_libint.cc:804-862,2119-2203 and
 * iface.cc). Only the 2-, 3- and 4-center ERI tasks exist; every other task
 * resolves to the (empty) `default` task as in engine.impl.h:592-599.
 */
#ifndef _libint2_oracle_params_h_
#define _libint2_oracle_params_h_
#define LIBINT2_API_PREFIX
#define LIBINT2_MAX_VECLEN 1
#define LIBINT2_ALIGN_SIZE 0
#define LIBINT2_REALTYPE double
#define LIBINT2_FLOP_COUNT 0
#define LIBINT2_CONTRACTED_INTS 1
#define LIBINT2_USE_COMPOSITE_EVALUATORS 0
#define LIBINT2_CARTGAUSS_MAX_AM 12
#define LIBINT2_CGSHELL_ORDERING 1
#define LIBINT2_CGSHELL_ORDERING_STANDARD 1
#define LIBINT2_CGSHELL_ORDERING_INTV3 2
#define LIBINT2_CGSHELL_ORDERING_GAMESS 3
#define LIBINT2_CGSHELL_ORDERING_ORCA 4
#define LIBINT2_CGSHELL_ORDERING_BAGEL 5
#define LIBINT2_SHELLQUARTET_SET 1
#define LIBINT2_SHELLQUARTET_SET_STANDARD 1
#define LIBINT2_SHELLQUARTET_SET_ORCA 2
#define LIBINT2_MAX_AM 6
#define LIBINT2_MAX_AM_default 6
#define LIBINT2_MAX_AM_eri 6
#define LIBINT2_MAX_AM_3eri 6
#define LIBINT2_MAX_AM_2eri 6
#define LIBINT2_SUPPORT_ERI 1
#define LIBINT2_DERIV_ERI_ORDER 0
#define LIBINT2_SUPPORT_ERI3 1
#define LIBINT2_DERIV_ERI3_ORDER 0
#define LIBINT2_SUPPORT_ERI2 1
#define LIBINT2_DERIV_ERI2_ORDER 0
#define LIBINT2_MAX_DERIV_ORDER 0

#define LIBINT2_TASK_EXISTS_0overlap 0
#define LIBINT2_TASK_EXISTS_0kinetic 0
#define LIBINT2_TASK_EXISTS_0elecpot 0
#define LIBINT2_TASK_EXISTS_01emultipole 0
#define LIBINT2_TASK_EXISTS_02emultipole 0
#define LIBINT2_TASK_EXISTS_03emultipole 0
#define LIBINT2_TASK_EXISTS_0sphemultipole 0
#define LIBINT2_TASK_EXISTS_0opVop 0
#define LIBINT2_TASK_EXISTS_0eri 0
#define LIBINT2_TASK_EXISTS_0r12kg12 0
#define LIBINT2_TASK_EXISTS_0r12_0_g12 0
#define LIBINT2_TASK_EXISTS_0r12_2_g12 0
#define LIBINT2_TASK_EXISTS_0g12_T1_g12 0
#define LIBINT2_TASK_EXISTS_0g12dkh 0
#define LIBINT2_TASK_EXISTS_1overlap 0
#define LIBINT2_TASK_EXISTS_1kinetic 0
#define LIBINT2_TASK_EXISTS_1elecpot 0
#define LIBINT2_TASK_EXISTS_11emultipole 0
#define LIBINT2_TASK_EXISTS_12emultipole 0
#define LIBINT2_TASK_EXISTS_13emultipole 0
#define LIBINT2_TASK_EXISTS_1sphemultipole 0
#define LIBINT2_TASK_EXISTS_1opVop 0
#define LIBINT2_TASK_EXISTS_1eri 0
#define LIBINT2_TASK_EXISTS_1r12kg12 0
#define LIBINT2_TASK_EXISTS_1r12_0_g12 0
#define LIBINT2_TASK_EXISTS_1r12_2_g12 0
#define LIBINT2_TASK_EXISTS_1g12_T1_g12 0
#define LIBINT2_TASK_EXISTS_1g12dkh 0
#define LIBINT2_TASK_EXISTS_2overlap 0
#define LIBINT2_TASK_EXISTS_2kinetic 0
#define LIBINT2_TASK_EXISTS_2elecpot 0
#define LIBINT2_TASK_EXISTS_21emultipole 0
#define LIBINT2_TASK_EXISTS_22emultipole 0
#define LIBINT2_TASK_EXISTS_23emultipole 0
#define LIBINT2_TASK_EXISTS_2sphemultipole 0
#define LIBINT2_TASK_EXISTS_2opVop 0
#define LIBINT2_TASK_EXISTS_2eri 1
#define LIBINT2_TASK_EXISTS_2r12kg12 0
#define LIBINT2_TASK_EXISTS_2r12_0_g12 0
#define LIBINT2_TASK_EXISTS_2r12_2_g12 0
#define LIBINT2_TASK_EXISTS_2g12_T1_g12 0
#define LIBINT2_TASK_EXISTS_2g12dkh 0
#define LIBINT2_TASK_EXISTS_3overlap 0
#define LIBINT2_TASK_EXISTS_3kinetic 0
#define LIBINT2_TASK_EXISTS_3elecpot 0
#define LIBINT2_TASK_EXISTS_31emultipole 0
#define LIBINT2_TASK_EXISTS_32emultipole 0
#define LIBINT2_TASK_EXISTS_33emultipole 0
#define LIBINT2_TASK_EXISTS_3sphemultipole 0
#define LIBINT2_TASK_EXISTS_3opVop 0
#define LIBINT2_TASK_EXISTS_3eri 1
#define LIBINT2_TASK_EXISTS_3r12kg12 0
#define LIBINT2_TASK_EXISTS_3r12_0_g12 0
#define LIBINT2_TASK_EXISTS_3r12_2_g12 0
#define LIBINT2_TASK_EXISTS_3g12_T1_g12 0
#define LIBINT2_TASK_EXISTS_3g12dkh 0
#define LIBINT2_TASK_EXISTS_4overlap 0
#define LIBINT2_TASK_EXISTS_4kinetic 0
#define LIBINT2_TASK_EXISTS_4elecpot 0
#define LIBINT2_TASK_EXISTS_41emultipole 0
#define LIBINT2_TASK_EXISTS_42emultipole 0
#define LIBINT2_TASK_EXISTS_43emultipole 0
#define LIBINT2_TASK_EXISTS_4sphemultipole 0
#define LIBINT2_TASK_EXISTS_4opVop 0
#define LIBINT2_TASK_EXISTS_4eri 1
#define LIBINT2_TASK_EXISTS_4r12kg12 0
#define LIBINT2_TASK_EXISTS_4r12_0_g12 0
#define LIBINT2_TASK_EXISTS_4r12_2_g12 0
#define LIBINT2_TASK_EXISTS_4g12_T1_g12 0
#define LIBINT2_TASK_EXISTS_4g12dkh 0
#endif
