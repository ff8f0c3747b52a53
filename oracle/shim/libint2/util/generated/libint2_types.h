/* oracle/shim/libint2/util/generated/libint2_types.h -- TEST INFRASTRUCTURE (oracle).
 *
 * Stand-in for the generator-emitted evaluator type. Layout rules follow
 * /root/reference/src/bin/libint/iface.cc:485-683: one `double sym[VECLEN]`
 * per symbol, then stack / vstack / targets / veclen / contrdepth; under C++
 * stack, vstack and targets are `mutable` (context.cc:573-575) because build
 * functions take `const Libint_t*` yet set targets[0].
 * Only the symbols Engine::compute2 touches for Operator::coulomb, deriv 0
 * (engine.impl.h:1514-1701) are present; each one has LIBINT2_DEFINED_<sym>.
 */
#ifndef _libint2_oracle_types_h_
#define _libint2_oracle_types_h_
#include <libint2/util/generated/libint2_params.h>

#define LB200_ORACLE_MAX_M 24 /* 4*LIBINT2_MAX_AM */

#ifdef __cplusplus
#define LB200_MUTABLE mutable
#else
#define LB200_MUTABLE
#endif

#define LB200_SYM(s) double s[LIBINT2_MAX_VECLEN];

typedef struct {
  /* (ss|ss)^(m), m = 0..24; names per include/libint2.h:24-25 */
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_0)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_1)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_2)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_3)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_4)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_5)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_6)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_7)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_8)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_9)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_10)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_11)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_12)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_13)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_14)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_15)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_16)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_17)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_18)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_19)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_20)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_21)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_22)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_23)
  LB200_SYM(_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_24)
  LB200_SYM(PA_x) LB200_SYM(PA_y) LB200_SYM(PA_z)
  LB200_SYM(PB_x) LB200_SYM(PB_y) LB200_SYM(PB_z)
  LB200_SYM(QC_x) LB200_SYM(QC_y) LB200_SYM(QC_z)
  LB200_SYM(QD_x) LB200_SYM(QD_y) LB200_SYM(QD_z)
  LB200_SYM(AB_x) LB200_SYM(AB_y) LB200_SYM(AB_z)
  LB200_SYM(BA_x) LB200_SYM(BA_y) LB200_SYM(BA_z)
  LB200_SYM(CD_x) LB200_SYM(CD_y) LB200_SYM(CD_z)
  LB200_SYM(DC_x) LB200_SYM(DC_y) LB200_SYM(DC_z)
  LB200_SYM(WP_x) LB200_SYM(WP_y) LB200_SYM(WP_z)
  LB200_SYM(WQ_x) LB200_SYM(WQ_y) LB200_SYM(WQ_z)
  LB200_SYM(oo2z) LB200_SYM(oo2e) LB200_SYM(oo2ze) LB200_SYM(roz) LB200_SYM(roe)
  /* referenced (unguarded) by the 1-body branch of engine.impl.h:287-289,1054-1056
     and include/libint2.h:36; never used by the 2-body oracle */
  LB200_SYM(_0_Overlap_0_x) LB200_SYM(_0_Overlap_0_y) LB200_SYM(_0_Overlap_0_z)
  LB200_SYM(_aB_s___0___ElecPot_s___0___Ab__up_0)
  LB200_MUTABLE double* stack;
  LB200_MUTABLE double* vstack;
  LB200_MUTABLE double* targets[1];
  int veclen;
  int contrdepth;
} Libint_t;

#define LIBINT2_DEFINED__aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_0 1
#define LIBINT2_DEFINED_PA_x 1
#define LIBINT2_DEFINED_PA_y 1
#define LIBINT2_DEFINED_PA_z 1
#define LIBINT2_DEFINED_PB_x 1
#define LIBINT2_DEFINED_PB_y 1
#define LIBINT2_DEFINED_PB_z 1
#define LIBINT2_DEFINED_QC_x 1
#define LIBINT2_DEFINED_QC_y 1
#define LIBINT2_DEFINED_QC_z 1
#define LIBINT2_DEFINED_QD_x 1
#define LIBINT2_DEFINED_QD_y 1
#define LIBINT2_DEFINED_QD_z 1
#define LIBINT2_DEFINED_AB_x 1
#define LIBINT2_DEFINED_AB_y 1
#define LIBINT2_DEFINED_AB_z 1
#define LIBINT2_DEFINED_BA_x 1
#define LIBINT2_DEFINED_BA_y 1
#define LIBINT2_DEFINED_BA_z 1
#define LIBINT2_DEFINED_CD_x 1
#define LIBINT2_DEFINED_CD_y 1
#define LIBINT2_DEFINED_CD_z 1
#define LIBINT2_DEFINED_DC_x 1
#define LIBINT2_DEFINED_DC_y 1
#define LIBINT2_DEFINED_DC_z 1
#define LIBINT2_DEFINED_WP_x 1
#define LIBINT2_DEFINED_WP_y 1
#define LIBINT2_DEFINED_WP_z 1
#define LIBINT2_DEFINED_WQ_x 1
#define LIBINT2_DEFINED_WQ_y 1
#define LIBINT2_DEFINED_WQ_z 1
#define LIBINT2_DEFINED_oo2z 1
#define LIBINT2_DEFINED_oo2e 1
#define LIBINT2_DEFINED_oo2ze 1
#define LIBINT2_DEFINED_roz 1
#define LIBINT2_DEFINED_roe 1

#endif
