/*
 * acgb200/config.h -- build configuration of the B200-native aCG hot path.
 *
 * ABI counterpart of the reference's acg/config.h:30-96.  Only the default
 * index width of the reference (acgidx_t == int, ACG_IDX_SIZE undefined) is
 * supported: device column indices are 32-bit, which is what keeps the CSR
 * stream at 12 B per nonzero.
 */
#ifndef ACGB200_CONFIG_H
#define ACGB200_CONFIG_H

#include <inttypes.h>
#include <limits.h>
#include <stdint.h>

#if defined(ACG_IDX_SIZE) && ACG_IDX_SIZE != 32
#error "acgb200 supports 32-bit acgidx_t only (reference default, acg/config.h:61-70)"
#endif

typedef int acgidx_t;          /* acg/config.h:62 */
#define PRIdx "d"
#define ACGIDX_T_MIN INT_MIN
#define ACGIDX_T_MAX INT_MAX

#ifndef ACG_API
#define ACG_API __attribute__((visibility("default")))
#endif

#endif
