#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include "R.h"
typedef struct SEXPREC* SEXP;
typedef unsigned int SEXPTYPE;
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
typedef enum { FALSE = 0, TRUE } Rboolean;
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
extern double R_NaReal, R_NaN;
#define NA_REAL R_NaReal
#define ISNAN(x) ((x) != (x))
int TYPEOF(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_alloc3DArray(SEXPTYPE, int, int, int);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_mkChar(const char*);
SEXP Rf_mkNamed(SEXPTYPE, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_length(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char* R_CHAR(SEXP);
#define CHAR(x) R_CHAR(x)
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
#endif
