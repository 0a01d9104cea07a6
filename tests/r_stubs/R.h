/* Minimal stand-in for R's <R.h>: declarations only, enough to type-check src/ppcseq_b200_shim.c where R is not
 * installed (tests/test_shim_compiles.py).  Signatures follow R 4.x's public C API. */
#ifndef PPCSEQ_STUB_R_H
#define PPCSEQ_STUB_R_H
#include <stddef.h>
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
void Rf_warning(const char *fmt, ...);
void R_CheckUserInterrupt(void);
#endif
