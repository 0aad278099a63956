// pcfd_comm.cuh -- collective-free halo exchange between one-process-per-GPU ranks (included by pcfd_kernels.cu).
//
// Replaces PObj::UpdateGeneralVectors (ucs/parallel.tcc:779-873: per-call index resend, pack loop, MPI_Isend/Irecv,
// MPI_Waitall) by DIRECT PUTS over NVLink: every rank maps its peers' field allocations and one small "flag page"
// through CUDA IPC once (pcfd_comm_export / pcfd_comm_connect; the host moves the fixed-size blobs with whatever
// transport it has -- MPI_Allgather in ucs.x, torch.distributed in the tests).  An exchange of a field is then
//
//   post:  ONE kernel (k_comm_put) that (1) tells every neighbour "I have reached exchange #e" (ready flag: all my
//          kernels that read the ghost rows about to be overwritten are behind me in stream order), (2) waits until
//          the neighbours it writes to have said the same, (3) gathers the rows each neighbour wants (the
//          reference's nodePackingList, peer's ghost order) and stores them straight into the neighbour's ghost
//          segment, (4) fences and -- last block out -- raises the neighbour's done flag to e;
//   wait:  a one-block kernel that spins until every neighbour it receives from has raised done >= e.
//
// No collective, no host synchronisation, no staging buffer; post and wait are separate calls so that
// ghost-independent work (the interior-edge flux kernel, the gather kernels of nodes without ghost neighbours) runs
// between them while the rows are in flight.  Flags are 64-bit epoch counters (monotonic, so a late
// reader can never miss one), written with st.release.sys after __threadfence_system() and read with ld.acquire.sys.
// Every spin has a wall-clock limit (globaltimer): a lost peer raises the error word instead of hanging the GPU.
//
// The same flag page carries a small all-gather (<= 8 doubles per rank, double-buffered by epoch parity) used for the
// global pressure-clip decision and for norms / minima -- the reference's three tiny MPI_Allreduce calls.
#pragma once

#include <stdint.h>
#include <unistd.h>

namespace {

constexpr int COMM_MAXR = PCFD_COMM_MAX_RANKS;
constexpr int COMM_GW = 8;                      // doubles per rank in the small all-gather
constexpr unsigned COMM_MAGIC = 0x70636664u;    // "pcfd"
constexpr unsigned long long COMM_SPIN_LIMIT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // PCFD_COMM_SPIN_SECONDS overrides

// layout of a rank's flag page (device memory, written by the peers)
struct CommFlags {
  unsigned long long ready[COMM_MAXR];          // ready[p]: peer p has reached exchange epoch ...
  unsigned long long done[COMM_MAXR];           // done[p]:  peer p's puts of epoch ... have landed here
  unsigned long long gtag[2][COMM_MAXR];        // small all-gather: epoch tag per (parity, rank)
  double gval[2][COMM_MAXR][COMM_GW];
  int err;                                      // raised by a spin that ran into the time limit
  int pad;
};

// what a rank publishes
struct CommBlob {
  unsigned magic;
  int rank, nranks, nnode, gnode, nbn, device, pid;
  unsigned long long token;                     // identifies the process (same pid + token: raw pointers are usable)
  int recv_offsets[COMM_MAXR + 1];
  unsigned char present[PCFD_F_COUNT];
  cudaIpcMemHandle_t fields[PCFD_F_COUNT];
  cudaIpcMemHandle_t flags;
  void* raw_fields[PCFD_F_COUNT];
  void* raw_flags;
};

// per-rank device tables read by the kernels
struct CommTable {
  int npeers_send, npeers_recv, nranks, me;
  unsigned long long spin_limit_ns;
  int send_peer[COMM_MAXR];                     // ranks this rank sends rows to
  int send_off[COMM_MAXR + 1];                  // prefix of their row counts in send_list
  int recv_peer[COMM_MAXR];                     // ranks this rank receives rows from
  CommFlags* peer_flags[COMM_MAXR];             // by RANK: mapped flag page of every rank (own page at [me])
  double* dst[PCFD_F_COUNT][COMM_MAXR];         // by field, send-peer slot: where my rows start in that peer's ghost segment
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want; false (and the error word raised) after COMM_SPIN_LIMIT_NS
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long want, int* err,
                                           unsigned long long limit_ns) {
  if (ld_acquire_sys(p) >= want) return true;
  const unsigned long long t0 = global_ns();
  unsigned backoff = 32;
  while (ld_acquire_sys(p) < want) {
    __nanosleep(backoff);
    if (backoff < 1024) backoff *= 2;
    if (global_ns() - t0 > limit_ns) { *err = 1; return false; }
  }
  return true;
}

// post of one or two fields: ready handshake, gather + put, done flags.  counter: zeroed int in local memory.
struct PutField { int field, n; const double* v; };
__global__ void __launch_bounds__(256) k_comm_put(const CommTable* __restrict__ tb, PutField f0, PutField f1,
                                                  const int* __restrict__ list, unsigned long long epoch, int* counter) {
  __shared__ int s_last;
  const int me = tb->me, np = tb->npeers_send;
  CommFlags* mine = tb->peer_flags[me];
  if (blockIdx.x == 0 && threadIdx.x < tb->npeers_recv) {
    // the ranks that write into MY ghost rows learn that I am done reading the old ones
    st_release_sys(&tb->peer_flags[tb->recv_peer[threadIdx.x]]->ready[me], epoch);
  }
  if (threadIdx.x < np) spin_until(&mine->ready[tb->send_peer[threadIdx.x]], epoch, &mine->err, tb->spin_limit_ns);
  __syncthreads();
  const int rows = tb->send_off[np];
  const long long total0 = (long long)rows * f0.n, total = total0 + (long long)rows * f1.n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const bool second = t >= total0;
    const long long tt = second ? t - total0 : t;
    const int n = second ? f1.n : f0.n;
    const int j = (int)(tt / n), cidx = (int)(tt - (long long)j * n);
    int s = 0;
    while (j >= tb->send_off[s + 1]) s++;
    const double* v = second ? f1.v : f0.v;
    tb->dst[second ? f1.field : f0.field][s][(size_t)(j - tb->send_off[s]) * n + cidx] = v[(size_t)list[j] * n + cidx];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < np) st_release_sys(&tb->peer_flags[tb->send_peer[threadIdx.x]]->done[me], epoch);
    if (threadIdx.x == 0) *counter = 0;
  }
}

__global__ void k_comm_wait(const CommTable* __restrict__ tb, unsigned long long epoch) {
  CommFlags* mine = tb->peer_flags[tb->me];
  if (threadIdx.x < tb->npeers_recv) spin_until(&mine->done[tb->recv_peer[threadIdx.x]], epoch, &mine->err, tb->spin_limit_ns);
}

// small all-gather: every rank writes its n values (doubles, or one int flag) into slot [parity][me] of every rank
__global__ void k_comm_bcast(const CommTable* __restrict__ tb, const double* __restrict__ vals, const int* __restrict__ iflag,
                             int n, unsigned long long gepoch) {
  const int r = threadIdx.x, me = tb->me, par = (int)(gepoch & 1ull);
  if (r >= tb->nranks) return;
  CommFlags* f = tb->peer_flags[r];
  for (int k = 0; k < n; k++) f->gval[par][me][k] = iflag ? (double)iflag[k] : vals[k];
  __threadfence_system();
  st_release_sys(&f->gtag[par][me], gepoch);
}
__global__ void k_comm_gwait(const CommTable* __restrict__ tb, int n, unsigned long long gepoch, double* __restrict__ out) {
  const int r = threadIdx.x, par = (int)(gepoch & 1ull);
  if (r >= tb->nranks) return;
  CommFlags* mine = tb->peer_flags[tb->me];
  spin_until(&mine->gtag[par][r], gepoch, &mine->err, tb->spin_limit_ns);
  for (int k = 0; k < n; k++) out[r * COMM_GW + k] = mine->gval[par][r][k];
}

}  // namespace

struct pcfd_comm {
  bool connected = false;
  // the halo of q at solutionSpace.tcc:665 repeats the one at :857 bit for bit when nothing wrote owned rows of q in
  // between: UpdateBCs only writes phantom rows unless a rank owns Dirichlet-type (hard-set) half-edges.  ghost_q_fresh:
  // the ghost rows of q hold what the owners hold (every rank runs the same call sequence, so the flag agrees across
  // ranks); no_hardset: no rank owns hard-set BC nodes (from the blobs).  PCFD_COMM_ELIDE=0 keeps every exchange.
  bool ghost_q_fresh = false, no_hardset = false, elide = true;
  long long elided = 0;
  CommFlags* flags = nullptr;                   // own flag page
  CommTable* table = nullptr;                   // device copy
  CommTable h{};                                // host copy
  int* counter = nullptr;
  double* gout = nullptr;                       // device [COMM_MAXR * COMM_GW]
  double* hgout = nullptr;                      // pinned host copy
  unsigned long long epoch = 0, gepoch = 0;
  unsigned long long field_epoch[PCFD_F_COUNT] = {};
  std::vector<void*> opened;                    // IPC mappings to close
  cudaEvent_t ev = nullptr;
};

namespace {

inline unsigned long long comm_token() {
  static int anchor;
  return (unsigned long long)(uintptr_t)&anchor ^ ((unsigned long long)getpid() << 32);
}

inline bool comm_on(const pcfd_ctx* c) { return c->comm && c->comm->connected; }

// field2 >= 0: a second field travels in the same kernel (one handshake, one epoch): qgrad + limiter
int comm_post(pcfd_ctx* c, int field, int field2 = -1) {
  pcfd_comm* m = c->comm;
  const int n = field_width(c, field);
  if (n == 0 || !c->f[field]) return fail(c, "pcfd_comm_post: field cannot be exchanged");
  const int n2 = field2 >= 0 ? field_width(c, field2) : 0;
  if (field2 >= 0 && (n2 == 0 || !c->f[field2])) return fail(c, "pcfd_comm_post: second field cannot be exchanged");
  m->epoch++;
  m->field_epoch[field] = m->epoch;
  if (field2 >= 0) m->field_epoch[field2] = m->epoch;
  if (field == PCFD_F_Q || field2 == PCFD_F_Q) {
    c->qmm_valid = false;   // ghost rows of q change: cached neighbour min / max are stale
    m->ghost_q_fresh = true;
  }
  const long long total = (long long)c->send_total * (n + n2);
  const int grid = (int)std::max<long long>(1, std::min<long long>((total + 1023) / 1024, (long long)c->num_sms));
  PROF("k_comm_put");
  k_comm_put<<<grid, 256, 0, c->stream>>>(m->table, PutField{field, n, c->f[field]},
                                          PutField{field2 >= 0 ? field2 : field, n2, field2 >= 0 ? c->f[field2] : nullptr},
                                          c->send_list, m->epoch, m->counter);
  LAUNCH_CHECK();
  return 0;
}

int comm_wait(pcfd_ctx* c, int field) {
  pcfd_comm* m = c->comm;
  if (m->field_epoch[field] == 0) return 0;   // nothing posted yet: nothing to wait for
  PROF("k_comm_wait");
  k_comm_wait<<<1, COMM_MAXR, 0, c->stream>>>(m->table, m->field_epoch[field]);
  LAUNCH_CHECK();
  return 0;
}

int comm_update(pcfd_ctx* c, int field) {
  if (comm_post(c, field)) return 1;
  return comm_wait(c, field);
}

// enqueue the all-gather of n doubles (device) or n ints (device flag words); the result lands in m->hgout
// ([rank][COMM_GW], pinned) once m->ev has fired
int comm_gather_enqueue(pcfd_ctx* c, const double* dvals, const int* dints, int n) {
  pcfd_comm* m = c->comm;
  if (n < 1 || n > COMM_GW) return fail(c, "pcfd_comm_allgather: 1..8 values per rank");
  m->gepoch++;
  PROF("k_comm_bcast");
  k_comm_bcast<<<1, COMM_MAXR, 0, c->stream>>>(m->table, dvals, dints, n, m->gepoch);
  LAUNCH_CHECK();
  PROF("k_comm_gwait");
  k_comm_gwait<<<1, COMM_MAXR, 0, c->stream>>>(m->table, n, m->gepoch, m->gout);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(m->hgout, m->gout, sizeof(double) * COMM_MAXR * COMM_GW, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaEventRecord(m->ev, c->stream));
  return 0;
}

// the error word, read through the context's own stream into pinned memory: a synchronous cudaMemcpy would go through
// the legacy default stream and can wait for kernels of OTHER contexts on the device (ranks as threads: a peer's put
// kernel that is spinning for this very rank)
int comm_check_err(pcfd_ctx* c) {
  pcfd_comm* m = c->comm;
  int* herr = reinterpret_cast<int*>(m->hgout + (size_t)COMM_MAXR * COMM_GW);
  CK(cudaMemcpyAsync(herr, &m->flags->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (*herr) return fail(c, "pcfd_comm: a peer did not answer within the time limit (flag spin timed out)");
  return 0;
}

}  // namespace

extern "C" {

size_t pcfd_comm_blob_size(void) { return sizeof(CommBlob); }

int pcfd_comm_export(pcfd_ctx* c, void* blob) {
  if (!c) return 1;
  if (!blob) return fail(c, "pcfd_comm_export: null blob");
  if (c->send_offsets.empty()) return fail(c, "pcfd_comm_export: pcfd_halo_configure has not been called");
  if (c->nranks > COMM_MAXR) return fail(c, "pcfd_comm_export: more ranks than PCFD_COMM_MAX_RANKS");
  CK(cudaSetDevice(c->device));
  if (!c->comm) {
    c->comm = new pcfd_comm();
    pcfd_comm* m = c->comm;
    if (dev_alloc(c, reinterpret_cast<char**>(&m->flags), 65536)) return 1;   // its own allocation: IPC maps whole allocations
    CK(cudaMemset(m->flags, 0, 65536));
    if (dev_alloc(c, &m->table, 1)) return 1;
    if (dev_alloc(c, &m->counter, 4)) return 1;
    CK(cudaMemset(m->counter, 0, 4 * sizeof(int)));
    if (dev_alloc(c, &m->gout, (size_t)COMM_MAXR * COMM_GW)) return 1;
    CK(cudaMallocHost(reinterpret_cast<void**>(&m->hgout), sizeof(double) * (COMM_MAXR * COMM_GW + 16)));   // + error word, upload slot
    CK(cudaEventCreateWithFlags(&m->ev, cudaEventDisableTiming));
  }
  CommBlob b;
  memset(&b, 0, sizeof(b));
  b.magic = COMM_MAGIC;
  b.rank = c->rank; b.nranks = c->nranks; b.nnode = c->nnode; b.gnode = c->gnode; b.nbn = c->nbn;
  b.device = c->device; b.pid = (int)getpid(); b.token = comm_token();
  for (int p = 0; p <= c->nranks; p++) b.recv_offsets[p] = c->recv_offsets[p];
  for (int k = 0; k < PCFD_F_COUNT; k++) {
    if (field_width(c, k) == 0 || !c->f[k] || c->fsize[k] == 0) continue;
    b.present[k] = 1;
    b.raw_fields[k] = c->f[k];
    CK(cudaIpcGetMemHandle(&b.fields[k], c->f[k]));
  }
  b.raw_flags = c->comm->flags;
  CK(cudaIpcGetMemHandle(&b.flags, c->comm->flags));
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int pcfd_comm_connect(pcfd_ctx* c, const void* blobs) {
  if (!c) return 1;
  if (!c->comm || !blobs) return fail(c, "pcfd_comm_connect: call pcfd_comm_export first and pass every rank's blob");
  CK(cudaSetDevice(c->device));
  pcfd_comm* m = c->comm;
  const CommBlob* B = static_cast<const CommBlob*>(blobs);
  const int R = c->nranks, me = c->rank;
  CommTable& t = m->h;
  memset(&t, 0, sizeof(t));
  t.nranks = R; t.me = me;
  t.spin_limit_ns = COMM_SPIN_LIMIT_NS;
  if (const char* e = getenv("PCFD_COMM_SPIN_SECONDS")) t.spin_limit_ns = (unsigned long long)(atof(e) * 1e9);
  auto open = [&](const CommBlob& b, const cudaIpcMemHandle_t& h, void* raw, void** out) -> int {
    if (b.pid == (int)getpid() && b.token == comm_token()) { *out = raw; return 0; }   // same process: plain pointer
    CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    m->opened.push_back(*out);
    return 0;
  };
  bool hardset = false;
  if (const char* e = getenv("PCFD_COMM_ELIDE")) m->elide = atoi(e) != 0;
  for (int r = 0; r < R; r++) {
    const CommBlob& b = B[r];
    if (b.magic != COMM_MAGIC || b.rank != r || b.nranks != R) return fail(c, "pcfd_comm_connect: blob table is not in rank order");
    hardset = hardset || b.nbn > 0;
    if (r == me) { t.peer_flags[r] = m->flags; continue; }
    void* p = nullptr;
    if (open(b, b.flags, b.raw_flags, &p)) return 1;
    t.peer_flags[r] = static_cast<CommFlags*>(p);
  }
  for (int r = 0; r < R; r++) {
    if (r == me) continue;
    if (c->recv_counts[r] > 0) t.recv_peer[t.npeers_recv++] = r;
    if (c->send_counts[r] == 0) continue;
    const int s = t.npeers_send++;
    t.send_peer[s] = r;
    t.send_off[s + 1] = t.send_off[s] + c->send_counts[r];
    const CommBlob& b = B[r];
    for (int k = 0; k < PCFD_F_COUNT; k++) {
      if (!b.present[k]) continue;
      void* base = nullptr;
      if (open(b, b.fields[k], b.raw_fields[k], &base)) return 1;
      const int n = field_width(c, k);
      // my rows land behind the peer's own nodes, at the peer's receive offset for me (parallel.tcc:848-864)
      t.dst[k][s] = static_cast<double*>(base) + ((size_t)b.nnode + b.recv_offsets[me]) * n;
    }
  }
  if (t.send_off[t.npeers_send] != c->send_total) return fail(c, "pcfd_comm_connect: send counts do not add up");
  CK(cudaMemcpy(m->table, &t, sizeof(t), cudaMemcpyHostToDevice));
  m->no_hardset = !hardset;
  m->ghost_q_fresh = false;
  m->connected = true;
  return 0;
}

int pcfd_comm_disconnect(pcfd_ctx* c) {
  if (!c || !c->comm) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (void* p : c->comm->opened) cudaIpcCloseMemHandle(p);
  c->comm->opened.clear();
  c->comm->connected = false;
  return 0;
}

int pcfd_comm_connected(const pcfd_ctx* c) { return c && comm_on(c) ? 1 : 0; }

/* diagnostics: ready[r], done[r] of this rank's flag page for r < nranks, then the error word (2*nranks + 1 values) */
int pcfd_comm_debug_flags(pcfd_ctx* c, unsigned long long* out) {
  if (!c || !c->comm || !out) return 1;
  CK(cudaSetDevice(c->device));
  CommFlags h;
  CK(cudaMemcpy(&h, c->comm->flags, sizeof(h), cudaMemcpyDeviceToHost));
  for (int r = 0; r < c->nranks; r++) { out[r] = h.ready[r]; out[c->nranks + r] = h.done[r]; }
  out[2 * c->nranks] = (unsigned long long)h.err;
  return 0;
}

int pcfd_comm_post(pcfd_ctx* c, int field) {
  if (!c) return 1;
  if (!comm_on(c)) return fail(c, "pcfd_comm_post: not connected");
  CK(cudaSetDevice(c->device));
  return comm_post(c, field);
}
int pcfd_comm_wait(pcfd_ctx* c, int field) {
  if (!c) return 1;
  if (!comm_on(c)) return fail(c, "pcfd_comm_wait: not connected");
  CK(cudaSetDevice(c->device));
  return comm_wait(c, field);
}
int pcfd_comm_update(pcfd_ctx* c, int field) {
  if (!c) return 1;
  if (!comm_on(c)) return fail(c, "pcfd_comm_update: not connected");
  CK(cudaSetDevice(c->device));
  return comm_update(c, field);
}

/* out[r*n + k] = value k of rank r; host buffers; synchronises the stream */
int pcfd_comm_allgather(pcfd_ctx* c, const double* vals, int n, double* out) {
  if (!c) return 1;
  if (!comm_on(c)) return fail(c, "pcfd_comm_allgather: not connected");
  if (!vals || !out || n < 1 || n > COMM_GW) return fail(c, "pcfd_comm_allgather: 1..8 values per rank");
  CK(cudaSetDevice(c->device));
  pcfd_comm* m = c->comm;
  double* stage = m->gout + (size_t)(COMM_MAXR - 1) * COMM_GW;   // tail of gout doubles as the upload slot
  if (c->nranks == COMM_MAXR) return fail(c, "pcfd_comm_allgather: staging slot unavailable at the maximum rank count");
  double* hstage = m->hgout + (size_t)COMM_MAXR * COMM_GW + 1;      // pinned: no staging through the driver's own buffers
  CK(cudaStreamSynchronize(c->stream));                             // (a previous gather's result may still be in flight)
  for (int k = 0; k < n; k++) hstage[k] = vals[k];
  CK(cudaMemcpyAsync(stage, hstage, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (comm_gather_enqueue(c, stage, nullptr, n)) return 1;
  CK(cudaEventSynchronize(m->ev));
  for (int r = 0; r < c->nranks; r++)
    for (int k = 0; k < n; k++) out[r * n + k] = m->hgout[r * COMM_GW + k];
  return comm_check_err(c);
}

}  // extern "C"
