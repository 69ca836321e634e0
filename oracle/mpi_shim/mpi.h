/*
 * Single-rank MPI stand-in -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * The reference's CPU solver (lanl-implementation/npts.c, test_npts.c, time_npts.c) is written
 * against MPI, which this image does not have.  With exactly one rank every collective it uses
 * degenerates to a memcpy or a no-op, so this header supplies those 21 entry points as static
 * inlines and lets the UNMODIFIED reference sources compile with plain gcc (oracle/Makefile).
 * Nothing here is shipped in the product library.
 *
 * A datatype handle is simply its size in bytes (MPI_DOUBLE == 8, MPI_INT == 4); derived types
 * created by Type_create_subarray/Type_create_resized collapse to their element type, which is
 * exact when the "subarray" is the whole array (always the case at one rank).
 */
#ifndef CFD_B200_MPI_SHIM_H
#define CFD_B200_MPI_SHIM_H

#include <string.h>
#include <sys/time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_ORDER_C 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_SUCCESS 0

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }

static inline int MPI_Cart_create(MPI_Comm c, int nd, const int *dims, const int *periods,
                                  int reorder, MPI_Comm *out)
{ (void)c; (void)nd; (void)dims; (void)periods; (void)reorder; *out = 0; return 0; }

static inline int MPI_Cart_coords(MPI_Comm c, int rank, int nd, int *coords)
{ (void)c; (void)rank; for (int i = 0; i < nd; i++) coords[i] = 0; return 0; }

static inline int MPI_Cart_get(MPI_Comm c, int nd, int *dims, int *periods, int *coords)
{ (void)c; for (int i = 0; i < nd; i++) { dims[i] = 1; periods[i] = 0; coords[i] = 0; } return 0; }

static inline int MPI_Type_create_subarray(int nd, const int *sizes, const int *subsizes,
                                           const int *starts, int order, MPI_Datatype old,
                                           MPI_Datatype *newt)
{ (void)nd; (void)sizes; (void)subsizes; (void)starts; (void)order; *newt = old; return 0; }

static inline int MPI_Type_create_resized(MPI_Datatype old, long lb, long extent, MPI_Datatype *newt)
{ (void)lb; (void)extent; *newt = old; return 0; }

static inline int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype *t) { (void)t; return 0; }

static inline int MPI_Gather(const void *s, int n, MPI_Datatype st, void *r, int rn,
                             MPI_Datatype rt, int root, MPI_Comm c)
{ (void)rn; (void)rt; (void)root; (void)c; memcpy(r, s, (size_t)n * (size_t)st); return 0; }

static inline int MPI_Gatherv(const void *s, int n, MPI_Datatype st, void *r, const int *counts,
                              const int *displs, MPI_Datatype rt, int root, MPI_Comm c)
{ (void)counts; (void)root; (void)c;
  memcpy((char *)r + (size_t)displs[0] * (size_t)rt, s, (size_t)n * (size_t)st); return 0; }

static inline int MPI_Scatterv(const void *s, const int *counts, const int *displs, MPI_Datatype st,
                               void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{ (void)counts; (void)root; (void)c;
  memcpy(r, (const char *)s + (size_t)displs[0] * (size_t)st, (size_t)rn * (size_t)rt); return 0; }

/* point-to-point: never reached with a peer at one rank */
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *q)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; *q = 0; return 0; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *q)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; *q = 0; return 0; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return 0; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return 0; }
static inline int MPI_Wait(MPI_Request *q, MPI_Status *s) { (void)q; (void)s; return 0; }

static inline double MPI_Wtime(void)
{ struct timeval tv; gettimeofday(&tv, 0); return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec; }

#endif
