// Device-side data layout of the Coulomb-ERI path (plain structs, shared by host and device).
#pragma once
#include <cstdint>

namespace lb200 {

constexpr int kMaxShellL = 4;         // s..g per shell
constexpr int kBoysOrder = 7;         // degree of the interpolating polynomial
constexpr int kBoysNInt = 819;        // intervals on [0,117)
constexpr double kBoysTmax = 117.0;   // above: asymptotic upward recursion
constexpr int kBoysTableMmax = 24;    // table rows per interval - 1
constexpr int kPrimCounters = 64;     // profiling counters (EriParams::prim_counter)

// One primitive pair of a shell pair: what ShellPair::PrimPairData holds in the reference
// (include/libint2/shell.h:1084-1092) plus gamma and c_a*c_b folded into K (engine.impl.h:1331-1367),
// computed once instead of per quartet.  64 bytes = half a cache line, 16-byte aligned: a record moves
// as four 128-bit loads / cp.async pieces and never straddles a line.  PA = P - A is formed in the
// kernel from the pair's PairGeom (the same subtraction the reference does per quartet,
// engine.impl.h:1514-1520).
struct alignas(16) PrimPair {
  double P[3];    // (alpha_a A + alpha_b B)/gamma  (shell.h:1186-1194)
  double Kc;      // sqrt(2) pi^(5/4) exp(-rho |AB|^2)/gamma * c_a * c_b   (shell.h:1241-1243)
  double gamma;   // alpha_a + alpha_b
  double oog;     // 1/gamma
  double ln_scr;  // primitive-pair screening value (shell.h:1162-1165,1214-1232,1288-1290)
  double nonsph;  // nonsph_screen_fac, ScreeningMethod::Conservative only (shell.h:1196-1212)
};
static_assert(sizeof(PrimPair) == 64, "PrimPair layout");

// Per shell pair: everything a quartet needs besides the primitive records, in one 64-byte line.
struct alignas(16) PairGeom {
  double A[3];    // centre of the first shell
  double AB[3];   // A - B
  int bf[2];      // first basis function of each shell
  int shell[2];   // shell indices (first, second)
};
static_assert(sizeof(PairGeom) == 64, "PairGeom layout");

// A block of shell pairs of one class (la >= lb, fixed purity), first shell = higher AM.
struct PairBlock {
  int npair;
  int la, lb, pure_a, pure_b;
  int unit_b;              // second shell is Shell::unit(): PA = 0 exactly (engine.impl.h:1515)
  const int* prim_off;     // [npair+1] offsets into prim
  const PrimPair* prim;    // concatenated primitive pairs (screened)
  const PairGeom* geom;    // [npair]
  const int* shell;        // [npair][2]  shell indices, again, as a dense array for the screening kernel
  const double* schwarz;   // [npair]     sqrt(max|(ab|ab)|)          (Fock build only)
  const int* gidx;         // [npair]     canonical pair index s1(s1+1)/2+s2 (Fock build only)
  int max_nprim;           // largest number of primitive pairs kept by any pair of the block
};

enum ScreeningMethod : int {  // values follow shell.h:1041-1059
  kScreenOriginal = 0x0001,
  kScreenConservative = 0x0010,
  kScreenSchwarz = 0x0100,
  kScreenSchwarzInf = 0x1000
};

enum EriMode : int { kModeStoreCart = 0, kModeStore = 1, kModeFock = 2, kModePrereq = 3 };

// kModePrereq: the caller supplies the per-primitive prerequisites itself -- the members of the
// reference's Libint_t that Engine::compute2 fills (engine.impl.h:1514-1641): (ss|ss)^(m) =
// F_m(T) * pfac, PA, QC, WP, WQ, oo2z, oo2e, oo2ze, roz, roe -- in the CALLER's bra/ket orientation;
// the kernel picks the side it needs.  This is what libint2_build_eri[la][lb][lc][ld](Libint_t*)
// receives (iface.cu).
struct PrereqRec {
  double F[kBoysTableMmax + 1];
  double PA[3], QC[3], WP[3], WQ[3];
  double oo2z, oo2e, oo2ze, roz, roe;
};

struct EriParams {
  PairBlock bra, ket;        // kernel-internal orientation: bra = "lane side"
  const int2* tasks;         // (bra pair, ket pair) indices (store modes); null with prod_nk > 0
  // implicit Cartesian-product task list (store modes): task t = (prod_b0 + t / prod_nk, prod_k0 + t % prod_nk)
  // in the CALLER's bra/ket orientation (swap_tasks applies afterwards); no task array is read
  unsigned prod_nk;
  int prod_b0, prod_k0;
  // Fock mode: one 16-byte record per surviving quartet, written by the screening kernel:
  // x = bra pair, y = ket pair | degeneracy code << 30 (deg = 1 << code, hartree-fock++.cc:1683-1687),
  // (z, w) = ln of the engine precision of this quartet (hartree-fock++.cc:1693-1695) as a double
  const int4* ftasks;
  const unsigned* ntasks_dev;  // if non-null, task count is read from device memory
  unsigned ntasks;
  int swap_tasks;             // 1: tasks are (ket pair, bra pair) in kernel orientation
  int uncontracted;           // 1: no pair of either block holds more than one primitive pair
  unsigned* work_counter;    // dynamic scheduling counter (zeroed by host)
  unsigned long long* prim_counter;  // profiling only (else null): [kPrimCounters] surviving primitive quartets
  const double* boys;        // [kBoysNInt][kBoysTableMmax+1][8]
  // primitive screening (engine.impl.h:1313-1314,1371-1386)
  int screening;
  double ln_precision;
  double precision;
  // store modes
  double* out;               // [ntasks][out_stride]
  long long out_stride;
  int transpose_out;         // 1: caller's bra is the kernel's ket -> write [cd][ab]
  // Fock mode (tests/hartree-fock/hartree-fock++.cc:1574-1772)
  const double* D;
  double* F;
  int nbf;
  const double* Dnorm;       // [nshell][nshell] inf-norms of shell blocks of D
  int nshell;
  double fock_precision;
  double ln_needed_engine_precision;
  double needed_engine_precision;
  // kModePrereq: records of task t are prereq[prereq_off[t] .. prereq_off[t+1]); geom[t] = AB[3], CD[3]
  const PrereqRec* prereq;
  const int* prereq_off;
  const double* prereq_geom;
  // sparse cart->pure tables (context.cu) for the run-time purity path of the Fock digestion
  const int* sph_rowptr;     // [(kMaxShellL+1)][2*kMaxShellL+2]
  const int* sph_col;
  const double* sph_val;
  const int* sph_base;       // [(kMaxShellL+1)]
};

}  // namespace lb200
