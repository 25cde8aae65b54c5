#ifndef R_STUB_INTERNALS_H
#define R_STUB_INTERNALS_H
#include "R.h"
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
#define NA_INTEGER (-2147483647 - 1)
extern SEXP R_NilValue, R_NamesSymbol;
double *REAL(SEXP);
int *INTEGER(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *CHAR(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_mkNamed(unsigned int, const char **);
SEXP Rf_ScalarReal(double);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_isNull(SEXP);
int Rf_ncols(SEXP);
int Rf_nrows(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean R_ToplevelExec(void (*fun)(void *), void *data);
#endif
