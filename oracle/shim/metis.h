/*
  oracle/shim/metis.h -- TEST INFRASTRUCTURE ONLY.
  Identity-ordering stand-in for METIS so that ParOptSparseCholesky.cpp (off the
  hot path; only needed to link the reference) compiles without METIS.
*/
#ifndef PCU_ORACLE_METIS_SHIM_H
#define PCU_ORACLE_METIS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int idx_t;
#define METIS_NOPTIONS 40
#define METIS_OPTION_NUMBERING 17
#define METIS_OK 1
int METIS_SetDefaultOptions(idx_t *options);
int METIS_NodeND(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                 idx_t *options, idx_t *perm, idx_t *iperm);
#ifdef __cplusplus
}
#endif
#endif
