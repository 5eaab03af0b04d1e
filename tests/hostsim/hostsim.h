/* hostsim.h -- TEST INFRASTRUCTURE ONLY (see cuda_mock.c) */
#ifndef ACGB200_HOSTSIM_H
#define ACGB200_HOSTSIM_H
#include <stddef.h>

struct simop { void (*fn)(void *); void *args; size_t size; };
struct simgraph { struct simop *ops; int n, cap; };
extern struct simgraph *hostsim_capturing;

/* execute now, or -- inside a stream capture -- record for cudaGraphLaunch */
int hostsim_run_or_record(void (*fn)(void *), const void *args, size_t size);
#endif
