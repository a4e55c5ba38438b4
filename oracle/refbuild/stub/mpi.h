/*
 * Single-process MPI shim (TEST INFRASTRUCTURE, not product code).
 *
 * The container has no MPI; the reference's hot-path sources only need an
 * MPI whose every communicator has exactly one rank (the 1x1 Grid oracle,
 * SURVEY.md section 8c).  Collectives degenerate to memcpy, self send/recv
 * goes through a small FIFO.  Implemented in mpi_shim.c.
 */
#ifndef ELB200_ORACLE_MPI_SHIM_H
#define ELB200_ORACLE_MPI_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Op;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Errhandler;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; int count_bytes; } MPI_Status;
typedef void MPI_User_function(void*, void*, int*, MPI_Datatype*);

#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_GROUP_NULL 0
#define MPI_GROUP_EMPTY 1
#define MPI_IN_PLACE ((void*)1)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_SIMILAR 2
#define MPI_UNEQUAL 3
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_ERRORS_ARE_FATAL 0
#define MPI_ERRORS_RETURN 1
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL 0

enum {
  MPI_DATATYPE_NULL = 0, MPI_CHAR, MPI_UNSIGNED_CHAR, MPI_BYTE, MPI_SHORT, MPI_INT, MPI_UNSIGNED,
  MPI_LONG, MPI_UNSIGNED_LONG, MPI_LONG_LONG_INT, MPI_UNSIGNED_LONG_LONG, MPI_FLOAT, MPI_DOUBLE,
  MPI_LONG_DOUBLE, MPI_COMPLEX, MPI_DOUBLE_COMPLEX, MPI_C_FLOAT_COMPLEX, MPI_C_DOUBLE_COMPLEX,
  MPI_FLOAT_INT, MPI_DOUBLE_INT, MPI_LONG_INT, MPI_2INT, MPI_SHIM_FIRST_DERIVED
};
#define MPI_LONG_LONG MPI_LONG_LONG_INT
enum { MPI_OP_NULL = 0, MPI_MAX, MPI_MIN, MPI_SUM, MPI_PROD, MPI_LAND, MPI_LOR, MPI_LXOR,
       MPI_BAND, MPI_BOR, MPI_BXOR, MPI_MAXLOC, MPI_MINLOC, MPI_SHIM_FIRST_USER_OP };

int MPI_Init(int*, char***);
int MPI_Init_thread(int*, char***, int, int*);
int MPI_Initialized(int*);
int MPI_Finalize(void);
int MPI_Finalized(int*);
int MPI_Query_thread(int*);
int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
int MPI_Error_string(int, char*, int*);

int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_dup(MPI_Comm, MPI_Comm*);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm*);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm*);
int MPI_Comm_free(MPI_Comm*);
int MPI_Comm_group(MPI_Comm, MPI_Group*);
int MPI_Comm_compare(MPI_Comm, MPI_Comm, int*);
MPI_Comm MPI_Comm_f2c(int);
int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler);
int MPI_Errhandler_set(MPI_Comm, MPI_Errhandler);
int MPI_Cart_create(MPI_Comm, int, const int*, const int*, int, MPI_Comm*);
int MPI_Cart_sub(MPI_Comm, const int*, MPI_Comm*);

int MPI_Group_rank(MPI_Group, int*);
int MPI_Group_size(MPI_Group, int*);
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_excl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_union(MPI_Group, MPI_Group, MPI_Group*);
int MPI_Group_difference(MPI_Group, MPI_Group, MPI_Group*);
int MPI_Group_free(MPI_Group*);
int MPI_Group_compare(MPI_Group, MPI_Group, int*);
int MPI_Group_translate_ranks(MPI_Group, int, const int*, MPI_Group, int*);

int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Ibcast(void*, int, MPI_Datatype, int, MPI_Comm, MPI_Request*);
int MPI_Gather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Igather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm, MPI_Request*);
int MPI_Gatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int, MPI_Comm);
int MPI_Scatter(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Scatterv(const void*, const int*, const int*, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Alltoall(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce_scatter(const void*, void*, const int*, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce_scatter_block(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Scan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);

int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Issend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Irsend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Sendrecv(const void*, int, MPI_Datatype, int, int, void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Sendrecv_replace(void*, int, MPI_Datatype, int, int, int, int, MPI_Comm, MPI_Status*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitall(int, MPI_Request*, MPI_Status*);
int MPI_Test(MPI_Request*, int*, MPI_Status*);
int MPI_Iprobe(int, int, MPI_Comm, int*, MPI_Status*);
int MPI_Get_count(const MPI_Status*, MPI_Datatype, int*);
int MPI_Get_address(const void*, MPI_Aint*);

int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype*);
int MPI_Type_create_struct(int, const int*, const MPI_Aint*, const MPI_Datatype*, MPI_Datatype*);
int MPI_Type_create_resized(MPI_Datatype, MPI_Aint, MPI_Aint, MPI_Datatype*);
int MPI_Type_commit(MPI_Datatype*);
int MPI_Type_free(MPI_Datatype*);
int MPI_Op_create(MPI_User_function*, int, MPI_Op*);
int MPI_Op_free(MPI_Op*);

#ifdef __cplusplus
}
#endif
#endif
