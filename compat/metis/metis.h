/*
 * metis.h -- the part of the public METIS 5.1 interface that the reference's acg/metis.c uses, for
 * an image whose only METIS is the static archive inside the CUDA toolkit
 * ($CUDA/targets/x86_64-linux/lib/libmetis_static.a, built with 64-bit idx_t and 32-bit real_t) and
 * which ships no header for it.  Written from the documented METIS 5.1 API (manual, section 5);
 * used only by tools/build_driver.sh to compile the UNMODIFIED reference sources with
 * -DACG_HAVE_METIS, so that `acg-cuda A.mtx` partitions its rows as the reference does
 * (METIS_PartGraphRecursive, acg/metis.c:225-346).  The library itself declares the two prototypes
 * it needs locally (acg_b200/csrc/metis_rows.c).
 */
#ifndef ACGB200_COMPAT_METIS_H
#define ACGB200_COMPAT_METIS_H

#include <inttypes.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define METIS_VER_MAJOR 5
#define METIS_VER_MINOR 1
#define METIS_VER_SUBMINOR 0

#define IDXTYPEWIDTH 64
#define REALTYPEWIDTH 32
typedef int64_t idx_t;
typedef float real_t;
#define IDX_MAX INT64_MAX
#define IDX_MIN INT64_MIN
#define PRIDX PRId64
#define SCIDX SCNd64
#define PRREAL "f"

#define METIS_NOPTIONS 40

/* return codes */
typedef enum { METIS_OK = 1, METIS_ERROR_INPUT = -2, METIS_ERROR_MEMORY = -3, METIS_ERROR = -4 } rstatus_et;

/* indices into the options array */
typedef enum {
    METIS_OPTION_PTYPE, METIS_OPTION_OBJTYPE, METIS_OPTION_CTYPE, METIS_OPTION_IPTYPE, METIS_OPTION_RTYPE,
    METIS_OPTION_DBGLVL, METIS_OPTION_NITER, METIS_OPTION_NCUTS, METIS_OPTION_SEED, METIS_OPTION_NO2HOP,
    METIS_OPTION_MINCONN, METIS_OPTION_CONTIG, METIS_OPTION_COMPRESS, METIS_OPTION_CCORDER, METIS_OPTION_PFACTOR,
    METIS_OPTION_NSEPS, METIS_OPTION_UFACTOR, METIS_OPTION_NUMBERING
} moptions_et;

typedef enum { METIS_PTYPE_RB, METIS_PTYPE_KWAY } mptype_et;
typedef enum { METIS_CTYPE_RM, METIS_CTYPE_SHEM } mctype_et;
typedef enum { METIS_IPTYPE_GROW, METIS_IPTYPE_RANDOM, METIS_IPTYPE_EDGE, METIS_IPTYPE_NODE, METIS_IPTYPE_METISRB } miptype_et;
typedef enum { METIS_RTYPE_FM, METIS_RTYPE_GREEDY, METIS_RTYPE_SEP2SIDED, METIS_RTYPE_SEP1SIDED } mrtype_et;
typedef enum { METIS_OBJTYPE_CUT, METIS_OBJTYPE_VOL, METIS_OBJTYPE_NODE } mobjtype_et;
typedef enum {
    METIS_DBG_INFO = 1, METIS_DBG_TIME = 2, METIS_DBG_COARSEN = 4, METIS_DBG_REFINE = 8, METIS_DBG_IPART = 16,
    METIS_DBG_MOVEINFO = 32, METIS_DBG_SEPINFO = 64, METIS_DBG_CONNINFO = 128, METIS_DBG_CONTIGINFO = 256,
    METIS_DBG_MEMORY = 2048
} mdbglvl_et;

int METIS_SetDefaultOptions(idx_t *options);
int METIS_PartGraphRecursive(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
                             idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
                             idx_t *edgecut, idx_t *part);
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
                        idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
                        idx_t *edgecut, idx_t *part);
int METIS_NodeND(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *options, idx_t *perm, idx_t *iperm);

#ifdef __cplusplus
}
#endif
#endif
