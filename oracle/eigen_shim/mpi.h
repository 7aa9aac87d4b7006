// TEST INFRASTRUCTURE (oracle) -- single-rank stand-in for the MPI calls on the reference's hot path
// (src/rSVD.cpp:15-16,49,52; src/PM.cpp:8-9,60,68).  With one rank, Gatherv is a copy and Bcast is a no-op.
//
// One test hook: the reference draws Omega from std::random_device (src/rSVD.cpp:26-28), so its output is
// not reproducible.  oracle_mpi_set_bcast_override(p, count) makes the NEXT MPI_Bcast whose element count
// equals `count` deliver p[0..count) instead -- i.e. "rank 0 held a host-supplied Omega" -- which lets the
// unmodified rSVD() run on an Omega chosen by the test.  The override is one-shot.
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H
#include <cstddef>
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_SUCCESS 0
inline const double*& oracle_mpi_override_ptr() { static const double* p = nullptr; return p; }
inline std::size_t& oracle_mpi_override_count() { static std::size_t n = 0; return n; }
inline void oracle_mpi_set_bcast_override(const double* p, std::size_t count) { oracle_mpi_override_ptr() = p; oracle_mpi_override_count() = count; }
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
#define MPI_INT 4
// point-to-point calls only appear on multi-rank branches (image_compression/src/image_com.cpp:387,400); with one rank they are unreachable
inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return MPI_SUCCESS; }
inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
inline int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype, void* rbuf, const int*, const int* displs, MPI_Datatype, int, MPI_Comm) {
  std::memcpy(static_cast<char*>(rbuf) + static_cast<std::size_t>(displs[0]) * sizeof(double), sbuf, static_cast<std::size_t>(scount) * sizeof(double));
  return MPI_SUCCESS;
}
inline int MPI_Bcast(void* buf, int count, MPI_Datatype, int, MPI_Comm) {
  if (oracle_mpi_override_ptr() && oracle_mpi_override_count() == static_cast<std::size_t>(count)) {
    std::memcpy(buf, oracle_mpi_override_ptr(), static_cast<std::size_t>(count) * sizeof(double));
    oracle_mpi_override_ptr() = nullptr;
  }
  return MPI_SUCCESS;
}
#endif
