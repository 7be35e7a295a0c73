// ctx_internal.h -- the few context internals other translation units need (frieda_ctx is opaque).
#pragma once
#include "../../include/frieda_b200.h"

extern "C" {
int frieda_ctx_device(const frieda_ctx *ctx);
int frieda_ctx_fail_arg(frieda_ctx *ctx, const char *msg);
int frieda_ctx_fail_cuda(frieda_ctx *ctx, int cuda_error, const char *what);
void frieda_ctx_count_launches(frieda_ctx *ctx, unsigned n);
void frieda_ctx_prof_begin(frieda_ctx *ctx, const char *name);
void frieda_ctx_prof_end(frieda_ctx *ctx);
}
