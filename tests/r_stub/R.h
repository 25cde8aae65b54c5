/* Minimal declarations of the R C API used by ldsr_b200/r/ldsr_b200_shim.c, ONLY so that the
 * shim can be syntax/type-checked in an image without R (tests/test_r_shim_syntax.py).
 * Not R's headers; nothing here is linked or executed. */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...);
void R_CheckUserInterrupt(void);
void Rf_onintr(void);
#endif
