/* Minimal stand-in for <Rinternals.h> (see R.h in this directory). */
#ifndef PPCSEQ_STUB_RINTERNALS_H
#define PPCSEQ_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue;
extern int R_NaInt;
extern double R_NaReal;
#define NA_LOGICAL R_NaInt
#define NA_INTEGER R_NaInt
#define NA_REAL R_NaReal
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isNull(SEXP);
SEXP Rf_install(const char *);
SEXP Rf_mkString(const char *);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarLogical(int);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean onexit);
#endif
