/*
 * mpi_shim.cpp -- process-based MPI shim (TEST INFRASTRUCTURE ONLY; see mpi.h).
 *
 * MPI_Init forks $PCFD_MPI_NP-1 children; every pair of ranks shares one
 * AF_UNIX stream socket.  Sends are eager/buffered (copied into an output
 * queue), receives are matched by (source, tag) in arrival order, and all
 * progress happens inside MPI_Wait / the blocking calls through poll().
 * Reductions are evaluated on rank 0 in rank order (deterministic).
 */
#include "mpi.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>
#include <fcntl.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

namespace {

int g_rank = 0, g_np = 1;
bool g_init = false;
bool g_finalizing = false;          // inside MPI_Finalize: a peer that is through its barrier may hang up
std::vector<char> g_closed;         // peers that have hung up during finalisation
std::vector<int> g_fd;            // g_fd[peer]
std::vector<pid_t> g_children;    // rank 0 only

struct Msg { int tag; std::vector<char> data; };
struct RecvReq { int src, tag; char* buf; size_t nbytes; bool done; bool is_send; bool live; };

std::vector<std::deque<char> > g_out;         // bytes queued to each peer
std::vector<std::vector<char> > g_in;         // partial inbound stream per peer
std::vector<std::deque<Msg> > g_unexpected;   // complete, unmatched messages per peer
std::vector<RecvReq> g_req;

size_t TypeSize(MPI_Datatype t)
{
  switch(t){
  case MPI_BYTE: case MPI_CHAR: return 1;
  case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
  case MPI_DOUBLE: case MPI_COMPLEX: case MPI_LONG: return 8;
  case MPI_DOUBLE_COMPLEX: return 16;
  case MPI_DOUBLE_INT: return 16;   // struct {double; int;} padded
  }
  fprintf(stderr, "mpi_shim: unknown datatype %d\n", t);
  abort();
}

void Die(const char* what)
{
  fprintf(stderr, "mpi_shim[rank %d]: %s (errno %d: %s)\n", g_rank, what, errno, strerror(errno));
  _exit(97);
}

// try to match a completed inbound message from `src` against posted receives
bool MatchPosted(int src, Msg& m)
{
  for(size_t i = 0; i < g_req.size(); i++){
    RecvReq& r = g_req[i];
    if(r.live && !r.is_send && !r.done && r.src == src && r.tag == m.tag){
      if(m.data.size() > r.nbytes){
	fprintf(stderr, "mpi_shim[rank %d]: message truncation from %d tag %d (%zu > %zu)\n",
		g_rank, src, m.tag, m.data.size(), r.nbytes);
	_exit(98);
      }
      if(!m.data.empty()) memcpy(r.buf, m.data.data(), m.data.size());
      r.done = true;
      return true;
    }
  }
  return false;
}

void ParseInbound(int src)
{
  std::vector<char>& in = g_in[src];
  size_t off = 0;
  while(in.size() - off >= 2*sizeof(long)){
    long hdr[2];
    memcpy(hdr, in.data() + off, sizeof(hdr));
    size_t need = sizeof(hdr) + (size_t)hdr[1];
    if(in.size() - off < need) break;
    Msg m;
    m.tag = (int)hdr[0];
    m.data.assign(in.begin() + off + sizeof(hdr), in.begin() + off + need);
    off += need;
    if(!MatchPosted(src, m)) g_unexpected[src].push_back(std::move(m));
  }
  if(off) in.erase(in.begin(), in.begin() + off);
}

// one round of non-blocking I/O; returns true if any byte moved.  If block is
// set, sleeps in poll() until some socket is ready.
bool Progress(bool block)
{
  if(g_np == 1) return false;
  std::vector<pollfd> pfds;
  std::vector<int> peers;
  for(int p = 0; p < g_np; p++){
    if(p == g_rank || (!g_closed.empty() && g_closed[p])) continue;
    pollfd pf; pf.fd = g_fd[p]; pf.events = POLLIN; pf.revents = 0;
    if(!g_out[p].empty()) pf.events |= POLLOUT;
    pfds.push_back(pf); peers.push_back(p);
  }
  int rc = poll(pfds.data(), pfds.size(), block ? 1000 : 0);
  if(rc < 0 && errno != EINTR) Die("poll");
  bool moved = false;
  static char buf[1 << 16];
  for(size_t k = 0; k < pfds.size(); k++){
    int p = peers[k];
    if(pfds[k].revents & POLLOUT){
      std::deque<char>& q = g_out[p];
      while(!q.empty()){
	size_t n = q.size() < sizeof(buf) ? q.size() : sizeof(buf);
	std::copy(q.begin(), q.begin() + n, buf);
	ssize_t w = write(g_fd[p], buf, n);
	if(w < 0){
	  if(errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
	  Die("write");
	}
	q.erase(q.begin(), q.begin() + w);
	moved = true;
	if((size_t)w < n) break;
      }
    }
    if(pfds[k].revents & (POLLIN | POLLHUP)){
      for(;;){
	ssize_t r = read(g_fd[p], buf, sizeof(buf));
	if(r < 0){
	  if(errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR) break;
	  Die("read");
	}
	if(r == 0){
	  if(pfds[k].revents & POLLHUP){
	    // a peer leaves MPI_Finalize as soon as ITS barrier is complete, possibly while this rank still polls for
	    // the others: everything it owed has been read above
	    if(g_finalizing){ g_closed[p] = 1; break; }
	    fprintf(stderr, "mpi_shim[rank %d]: peer %d hung up\n", g_rank, p);
	    _exit(99);
	  }
	  break;
	}
	g_in[p].insert(g_in[p].end(), buf, buf + r);
	moved = true;
	if((size_t)r < sizeof(buf)) break;
      }
      ParseInbound(p);
    }
  }
  return moved;
}

void FlushAll()
{
  for(;;){
    bool pending = false;
    for(int p = 0; p < g_np; p++) if(p != g_rank && !g_out[p].empty()) pending = true;
    if(!pending) break;
    Progress(true);
  }
}

int NewReq(const RecvReq& r)
{
  for(size_t i = 0; i < g_req.size(); i++){
    if(!g_req[i].live){ g_req[i] = r; g_req[i].live = true; return (int)i; }
  }
  g_req.push_back(r);
  g_req.back().live = true;
  return (int)g_req.size() - 1;
}

const int TAG_COLL = -7000;

template <class T>
void ReduceTyped(T* acc, const T* in, int count, MPI_Op op)
{
  for(int i = 0; i < count; i++){
    switch(op){
    case MPI_SUM: acc[i] = acc[i] + in[i]; break;
    case MPI_MAX: if(in[i] > acc[i]) acc[i] = in[i]; break;
    case MPI_MIN: if(in[i] < acc[i]) acc[i] = in[i]; break;
    default: fprintf(stderr, "mpi_shim: bad op\n"); abort();
    }
  }
}

struct DoubleInt { double v; int i; };

void ReduceInto(void* acc, const void* in, int count, MPI_Datatype t, MPI_Op op)
{
  if(t == MPI_DOUBLE_INT){
    DoubleInt* a = (DoubleInt*)acc; const DoubleInt* b = (const DoubleInt*)in;
    for(int i = 0; i < count; i++){
      if(op == MPI_MINLOC){ if(b[i].v < a[i].v || (b[i].v == a[i].v && b[i].i < a[i].i)) a[i] = b[i]; }
      else if(op == MPI_MAXLOC){ if(b[i].v > a[i].v || (b[i].v == a[i].v && b[i].i < a[i].i)) a[i] = b[i]; }
      else { fprintf(stderr, "mpi_shim: bad op for DOUBLE_INT\n"); abort(); }
    }
    return;
  }
  switch(t){
  case MPI_INT: ReduceTyped((int*)acc, (const int*)in, count, op); break;
  case MPI_UNSIGNED: ReduceTyped((unsigned*)acc, (const unsigned*)in, count, op); break;
  case MPI_LONG: ReduceTyped((long*)acc, (const long*)in, count, op); break;
  case MPI_FLOAT: ReduceTyped((float*)acc, (const float*)in, count, op); break;
  case MPI_DOUBLE: ReduceTyped((double*)acc, (const double*)in, count, op); break;
  case MPI_DOUBLE_COMPLEX:
    if(op != MPI_SUM){ fprintf(stderr, "mpi_shim: complex op\n"); abort(); }
    ReduceTyped((double*)acc, (const double*)in, 2*count, op); break;
  case MPI_COMPLEX:
    if(op != MPI_SUM){ fprintf(stderr, "mpi_shim: complex op\n"); abort(); }
    ReduceTyped((float*)acc, (const float*)in, 2*count, op); break;
  default: fprintf(stderr, "mpi_shim: reduce on datatype %d\n", t); abort();
  }
}

} // namespace

extern "C" {

int MPI_Init(int*, char***)
{
  if(g_init) return MPI_SUCCESS;
  g_init = true;
  const char* e = getenv("PCFD_MPI_NP");
  g_np = e ? atoi(e) : 1;
  if(g_np < 1) g_np = 1;
  g_rank = 0;
  g_fd.assign(g_np, -1);
  if(g_np > 1){
    // sp[i][j] = rank i's end of the (i,j) socket
    std::vector<std::vector<int> > sp(g_np, std::vector<int>(g_np, -1));
    for(int i = 0; i < g_np; i++){
      for(int j = i+1; j < g_np; j++){
	int s[2];
	if(socketpair(AF_UNIX, SOCK_STREAM, 0, s)) Die("socketpair");
	sp[i][j] = s[0]; sp[j][i] = s[1];
      }
    }
    fflush(stdout); fflush(stderr);
    for(int r = 1; r < g_np; r++){
      pid_t pid = fork();
      if(pid < 0) Die("fork");
      if(pid == 0){ g_rank = r; g_children.clear(); break; }
      g_children.push_back(pid);
    }
    for(int i = 0; i < g_np; i++){
      for(int j = 0; j < g_np; j++){
	if(i == j) continue;
	if(i == g_rank){
	  g_fd[j] = sp[i][j];
	  int fl = fcntl(g_fd[j], F_GETFL, 0);
	  fcntl(g_fd[j], F_SETFL, fl | O_NONBLOCK);
	  int sz = 4 << 20;
	  setsockopt(g_fd[j], SOL_SOCKET, SO_SNDBUF, &sz, sizeof(sz));
	  setsockopt(g_fd[j], SOL_SOCKET, SO_RCVBUF, &sz, sizeof(sz));
	}
	else close(sp[i][j]);
      }
    }
  }
  g_out.assign(g_np, std::deque<char>());
  g_in.assign(g_np, std::vector<char>());
  g_unexpected.assign(g_np, std::deque<Msg>());
  return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
  if(!g_init) return MPI_SUCCESS;
  if(g_np > 1){
    g_finalizing = true;
    g_closed.assign(g_np, 0);
    MPI_Barrier(MPI_COMM_WORLD);
    FlushAll();
  }
  fflush(stdout); fflush(stderr);
  if(g_rank == 0){
    for(size_t i = 0; i < g_children.size(); i++){
      int st; waitpid(g_children[i], &st, 0);
    }
    g_children.clear();
  }
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int code)
{
  fflush(stdout); fflush(stderr);
  _exit(code ? code : 1);
}

int MPI_Comm_rank(MPI_Comm, int* rank){ *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size){ *size = g_np; return MPI_SUCCESS; }

double MPI_Wtime(void)
{
  timeval tv; gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6*tv.tv_usec;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request* req)
{
  size_t nbytes = (size_t)count*TypeSize(t);
  RecvReq r; r.src = dest; r.tag = tag; r.buf = (char*)buf; r.nbytes = nbytes; r.done = true; r.is_send = true;
  if(dest == g_rank){
    Msg m; m.tag = tag; m.data.assign((const char*)buf, (const char*)buf + nbytes);
    if(!MatchPosted(dest, m)) g_unexpected[dest].push_back(std::move(m));
  }
  else{
    long hdr[2] = {tag, (long)nbytes};
    std::deque<char>& q = g_out[dest];
    q.insert(q.end(), (char*)hdr, (char*)hdr + sizeof(hdr));
    q.insert(q.end(), (const char*)buf, (const char*)buf + nbytes);
    Progress(false);
  }
  *req = NewReq(r);
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request* req)
{
  RecvReq r; r.src = src; r.tag = tag; r.buf = (char*)buf; r.nbytes = (size_t)count*TypeSize(t);
  r.done = false; r.is_send = false;
  // unexpected queue first (in arrival order)
  std::deque<Msg>& u = g_unexpected[src];
  for(std::deque<Msg>::iterator it = u.begin(); it != u.end(); ++it){
    if(it->tag == tag){
      if(it->data.size() > r.nbytes){ fprintf(stderr, "mpi_shim: truncation\n"); _exit(98); }
      if(!it->data.empty()) memcpy(buf, it->data.data(), it->data.size());
      r.done = true;
      u.erase(it);
      break;
    }
  }
  *req = NewReq(r);
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* req, MPI_Status* status)
{
  if(*req < 0) return MPI_SUCCESS;
  while(!g_req[*req].done){
    if(!Progress(false)) Progress(true);
  }
  if(status){ status->MPI_SOURCE = g_req[*req].src; status->MPI_TAG = g_req[*req].tag; status->MPI_ERROR = 0; }
  g_req[*req].live = false;
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}

int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{
  MPI_Request r; MPI_Isend(buf, count, t, dest, tag, c, &r); return MPI_Wait(&r, NULL);
}

int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st)
{
  MPI_Request r; MPI_Irecv(buf, count, t, src, tag, c, &r); return MPI_Wait(&r, st);
}

int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm c)
{
  if(g_np == 1) return MPI_SUCCESS;
  if(g_rank == root){
    for(int p = 0; p < g_np; p++) if(p != root) MPI_Send(buf, count, t, p, TAG_COLL-1, c);
  }
  else MPI_Recv(buf, count, t, root, TAG_COLL-1, c, NULL);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
  size_t nbytes = (size_t)count*TypeSize(t);
  if(g_rank == root){
    std::vector<char> mine(nbytes), tmp(nbytes), acc(nbytes);
    memcpy(mine.data(), sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, nbytes);
    for(int p = 0; p < g_np; p++){
      const char* src = mine.data();
      if(p != root){ MPI_Recv(tmp.data(), count, t, p, TAG_COLL-2, c, NULL); src = tmp.data(); }
      if(p == 0) memcpy(acc.data(), src, nbytes);
      else ReduceInto(acc.data(), src, count, t, op);
    }
    memcpy(recvbuf, acc.data(), nbytes);
  }
  else MPI_Send(sendbuf, count, t, root, TAG_COLL-2, c);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  size_t nbytes = (size_t)count*TypeSize(t);
  if(g_np == 1){
    if(sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, nbytes);
    return MPI_SUCCESS;
  }
  if(g_rank == 0) MPI_Reduce(sendbuf, recvbuf, count, t, op, 0, c);
  else MPI_Reduce(sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, recvbuf, count, t, op, 0, c);
  return MPI_Bcast(recvbuf, count, t, 0, c);
}

int MPI_Barrier(MPI_Comm c)
{
  int a = 1, b = 0;
  return MPI_Allreduce(&a, &b, 1, MPI_INT, MPI_SUM, c);
}

int MPI_Allgather(const void* sendbuf, int scount, MPI_Datatype st, void* recvbuf, int rcount, MPI_Datatype rt, MPI_Comm c)
{
  size_t rbytes = (size_t)rcount*TypeSize(rt);
  char* rb = (char*)recvbuf;
  if(sendbuf != MPI_IN_PLACE) memcpy(rb + g_rank*rbytes, sendbuf, (size_t)scount*TypeSize(st));
  for(int p = 0; p < g_np; p++) if(p != g_rank) MPI_Send(rb + g_rank*rbytes, rcount, rt, p, TAG_COLL-3, c);
  for(int p = 0; p < g_np; p++) if(p != g_rank) MPI_Recv(rb + p*rbytes, rcount, rt, p, TAG_COLL-3, c, NULL);
  return MPI_SUCCESS;
}

int MPI_Alltoall(const void* sendbuf, int scount, MPI_Datatype st, void* recvbuf, int rcount, MPI_Datatype rt, MPI_Comm c)
{
  size_t sbytes = (size_t)scount*TypeSize(st), rbytes = (size_t)rcount*TypeSize(rt);
  const char* sb = (const char*)sendbuf; char* rb = (char*)recvbuf;
  memcpy(rb + g_rank*rbytes, sb + g_rank*sbytes, sbytes);
  for(int p = 0; p < g_np; p++) if(p != g_rank) MPI_Send(sb + p*sbytes, scount, st, p, TAG_COLL-4, c);
  for(int p = 0; p < g_np; p++) if(p != g_rank) MPI_Recv(rb + p*rbytes, rcount, rt, p, TAG_COLL-4, c, NULL);
  return MPI_SUCCESS;
}

} // extern "C"
