// mpi.h for PANSLBM2 programs built with _USE_MPI_DEFINES against panslbm2_b200 — NOT an MPI implementation.
// It maps the handful of MPI-1 calls the reference's drivers and host-side utilities make (production/heatsink3D.cpp:28-36,
// 136, 236, 272, 302-305; src/utility/mma.h:256-585; src/utility/vtkxmlexport.h:26, 172-214) onto the NCCL communicator of
// libpanslbm_b200.so: one process per GPU, rank == PEid.  The halo exchange of Stream()/iStream() and of the filters never goes
// through here; it lives inside the library.
//
// Launch with tools/mpiexec_b200 -n <ranks> <program> <args> (sets RANK, LOCAL_RANK, WORLD_SIZE and a job id), or under any
// launcher that sets RANK/WORLD_SIZE/LOCAL_RANK (torchrun does).  Rank 0 draws NCCL's unique id and publishes it through a
// file in $PANSLBM_RDV_DIR (default /tmp); without those variables the program is a world of one.
#pragma once
#if __has_include("../../../include/panslbm_c.h")
#include "../../../include/panslbm_c.h"
#else
#include <panslbm_c.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DOUBLE 0
#define MPI_INT 1
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)

namespace panslbm_mpi {
    struct State { int rank = 0, size = 1; bool up = false; std::vector<pl_p2p_op> pending; };
    inline State& st() { static State s; return s; }
    inline void die(const char* what) { std::fprintf(stderr, "panslbm_b200 mpi shim: %s: %s\n", what, pl_last_error()); std::abort(); }
    inline int env_int(const char* a, const char* b, const char* c, int dflt) {
        for (const char* n : {a, b, c}) { if (!n) continue; const char* v = std::getenv(n); if (v && *v) return std::atoi(v); }
        return dflt;
    }
    inline size_t type_size(MPI_Datatype t) { return t == MPI_DOUBLE ? sizeof(double) : sizeof(int); }
}

inline int MPI_Init(int*, char***) {
    using namespace panslbm_mpi;
    State& s = st();
    s.rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", 0);
    s.size = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", 1);
    const int local = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", nullptr, s.rank);
    if (pl_device_count() > 0 && pl_set_device(local % pl_device_count())) die("pl_set_device");
    if (s.size > 1) {
        const char* dir = std::getenv("PANSLBM_RDV_DIR");
        const char* job = std::getenv("PANSLBM_JOB_ID");
        const char* port = std::getenv("MASTER_PORT");
        const std::string path = std::string(dir && *dir ? dir : "/tmp") + "/panslbm_nccl_" + (job && *job ? job : "0") + "_" + (port && *port ? port : "0") + ".id";
        char id[128];
        if (s.rank == 0) {
            if (pl_comm_unique_id(id)) die("pl_comm_unique_id");
            const std::string tmp = path + ".tmp";
            FILE* f = std::fopen(tmp.c_str(), "wb");
            if (!f || std::fwrite(id, 1, 128, f) != 128) die("cannot write the rendezvous file");
            std::fclose(f);
            std::rename(tmp.c_str(), path.c_str());
        } else {
            bool ok = false;
            for (int tries = 0; tries < 3000 && !ok; ++tries) {      // up to five minutes
                FILE* f = std::fopen(path.c_str(), "rb");
                if (f) { ok = std::fread(id, 1, 128, f) == 128; std::fclose(f); }
                if (!ok) std::this_thread::sleep_for(std::chrono::milliseconds(100));
            }
            if (!ok) die("rank 0 never published the NCCL id (rendezvous file)");
        }
        if (pl_comm_init(id, s.rank, s.size)) die("pl_comm_init");
        double one = 1.0;
        if (pl_comm_allreduce(&one, 1, 0)) die("first all-reduce");       // everybody has joined: the file can go
        if (s.rank == 0) std::remove(path.c_str());
    }
    s.up = true;
    return MPI_SUCCESS;
}
inline int MPI_Finalize() { plh_sync(); pl_comm_destroy(); panslbm_mpi::st().up = false; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* n) { *n = panslbm_mpi::st().size; return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = panslbm_mpi::st().rank; return MPI_SUCCESS; }
// Buffers may be the program's mirrored field arrays (MPI_Allreduce(MPI_IN_PLACE, dfds, ...), an Isend of a field): the staging
// copies below run inside the CUDA runtime, where a stale host copy cannot be fetched on demand — bring them up to date first.
inline int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm) {
    const size_t nbytes = (size_t)count*panslbm_mpi::type_size(type);
    if (send != MPI_IN_PLACE && send != recv) plh_host_acquire(send, nbytes, 0);
    plh_host_acquire(recv, nbytes, 1);
    if (send != MPI_IN_PLACE && send != recv) std::memcpy(recv, send, (size_t)count*panslbm_mpi::type_size(type));
    if (pl_comm_allreduce_v(recv, (size_t)count, type == MPI_DOUBLE ? 0 : 1, op)) panslbm_mpi::die("MPI_Allreduce");
    return MPI_SUCCESS;
}
inline int MPI_Barrier(MPI_Comm) { double one = 1.0; if (pl_comm_allreduce(&one, 1, 0)) panslbm_mpi::die("MPI_Barrier"); return MPI_SUCCESS; }
// every rank's block lands in its slot of a zeroed array which is then summed: the root (and everybody else) has the gather
inline int MPI_Gather(const void* send, int scount, MPI_Datatype stype, void* recv, int, MPI_Datatype, int root, MPI_Comm) {
    using namespace panslbm_mpi;
    const size_t bytes = (size_t)scount*type_size(stype);
    plh_host_acquire(send, bytes, 0);
    if (st().rank == root) plh_host_acquire(recv, bytes*(size_t)st().size, 1);
    std::vector<char> all(bytes*(size_t)st().size, 0);
    std::memcpy(all.data() + bytes*(size_t)st().rank, send, bytes);
    if (pl_comm_allreduce_v(all.data(), (size_t)scount*(size_t)st().size, stype == MPI_DOUBLE ? 0 : 1, 0)) die("MPI_Gather");
    if (st().rank == root) std::memcpy(recv, all.data(), all.size());
    return MPI_SUCCESS;
}
// tags are dropped: messages between a pair of ranks are matched in issue order, which is how the reference posts them
inline int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int, MPI_Comm, MPI_Request* req) {
    panslbm_mpi::st().pending.push_back(pl_p2p_op{const_cast<void*>(buf), (size_t)count*panslbm_mpi::type_size(type), dest, 1});
    if (req) *req = (int)panslbm_mpi::st().pending.size();
    return MPI_SUCCESS;
}
inline int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int, MPI_Comm, MPI_Request* req) {
    panslbm_mpi::st().pending.push_back(pl_p2p_op{buf, (size_t)count*panslbm_mpi::type_size(type), source, 0});
    if (req) *req = (int)panslbm_mpi::st().pending.size();
    return MPI_SUCCESS;
}
inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) {
    auto& p = panslbm_mpi::st().pending;
    for (auto& o : p) plh_host_acquire(o.host, o.bytes, o.is_send ? 0 : 1);
    if (!p.empty() && pl_comm_p2p(p.data(), (int)p.size())) panslbm_mpi::die("MPI_Waitall");
    p.clear();
    return MPI_SUCCESS;
}
