// Plain (non-CUDA) view of a basis for host-only translation units.
#pragma once
struct lb200_basis;
struct lb200_basis_view {
  int nshell;
  const int *l, *pure, *nprim, *off;
  const double *O, *alpha, *coeff;
};
lb200_basis_view lb200_view(const lb200_basis* bs);
