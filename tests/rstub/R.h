/* Minimal stand-in for R's C API: just enough declarations to type-check r/src/ssgpu_rcall.c with gcc
 * -fsyntax-only in an image without R (tests/test_abi.py).  Not R, not linkable. */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stddef.h>
typedef ptrdiff_t R_xlen_t;
double norm_rand(void);
void GetRNGstate(void);
void PutRNGstate(void);
char* R_alloc(size_t n, int size);
#endif
