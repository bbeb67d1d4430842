/* oracle/refshim/mpi.h -- TEST INFRASTRUCTURE (not product code).
 * The reference's serial MPI stub header (src/STUBS/mpi.h) predates two symbols
 * that comm.cpp / comm_brick.cpp / irregular.cpp use. This wrapper includes the
 * stub header where it lies under /root/reference and adds the two missing names,
 * so the reference sources compile unmodified. */
#ifndef ORACLE_REFSHIM_MPI_H
#define ORACLE_REFSHIM_MPI_H
#include_next "mpi.h"   /* -> $(REF)/STUBS/mpi.h via -I order */
#ifndef MPI_STATUS_IGNORE
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#endif
static inline int MPI_Request_free(MPI_Request *) { return 0; }
#endif
