/* oracle MPI stub (test infrastructure): single-rank no-ops for the two calls
 * the hot-path translation units make (hot_x_section.c:717,769,825). */
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_INT 2
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
#endif
