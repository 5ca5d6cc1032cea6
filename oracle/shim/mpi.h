// Stand-in for <mpi.h> (test infrastructure; the image has no MPI).
// The reference only needs the datatype handle type and one constant
// (reference: src/simulator-mpi/mpi_ext.hpp:29-32, SwapperMT.cpp:18).
#pragma once
typedef int MPI_Datatype;
#define MPI_DOUBLE_COMPLEX 1
