/* See mpi.h.  Single-node, process-per-rank MPI subset over POSIX shm.  One control segment per job
 * (created zero-filled by the launcher) holds a sense-reversing barrier, the window registry, the
 * all-reduce staging slots and the point-to-point mailboxes.  RMA windows are the segments handed out
 * by MPI_Alloc_mem, mapped by every peer at MPI_Win_create; MPI_Put is a memcpy into the target's
 * mapping (passive target: the target CPU is not involved, exactly like the reference expects). */
#define _GNU_SOURCE
#include "mpi.h"
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define MAXR 64
#define MAXWIN 8
#define MAXALLOC 16
#define AR_CHUNK ((size_t)4 << 20) /* bytes of all-reduce staging per rank */
#define MB_SLOTS 4
#define MB_BYTES ((size_t)2 << 20) /* largest two-sided message */

struct mbox { _Atomic int state; int tag; long nbytes; };
struct winrec { char name[MAXR][64]; long size[MAXR]; int disp_unit[MAXR]; };
struct ctl {
    _Atomic int bar_count;
    _Atomic int bar_gen;
    _Atomic int aborted;
    int pad;
    struct winrec win[MAXWIN];
    struct mbox mb[MAXR][MAXR][MB_SLOTS]; /* [dst][src][slot] */
};

static struct ctl *g_ctl;
static char *g_ar;   /* all-reduce staging: size * AR_CHUNK */
static char *g_mbd;  /* mailbox payloads: [dst][src][slot] * MB_BYTES */
static int g_rank = 0, g_size = 1, g_nwin = 0, g_nalloc = 0;
static char g_job[48] = "solo";
static struct { void *ptr; size_t size; char name[64]; } g_alloc[MAXALLOC];
static struct { char *peer[MAXR]; int disp_unit[MAXR]; size_t size[MAXR]; } g_win[MAXWIN];

static void die(const char *msg) {
    fprintf(stderr, "[egmpi rank %d] fatal: %s\n", g_rank, msg);
    if (g_ctl) atomic_store(&g_ctl->aborted, 1);
    _exit(86);
}

static inline void relax(unsigned *spins) {
    if (++*spins & 0x3ff) {
        __builtin_ia32_pause();
    } else {
        if (g_ctl && atomic_load_explicit(&g_ctl->aborted, memory_order_relaxed)) _exit(87);
        sched_yield();
    }
}

static size_t ctl_bytes(void) { return (sizeof(struct ctl) + 4095) & ~(size_t)4095; }

static void *map_shm(const char *name, size_t size, int create) {
    int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) die("shm_open failed");
    if (create && ftruncate(fd, (off_t)size) != 0) die("ftruncate failed");
    void *p = mmap(NULL, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) die("mmap failed");
    return p;
}

int MPI_Init(int *argc, char ***argv) {
    (void)argc; (void)argv;
    const char *j = getenv("EGMPI_JOB"), *r = getenv("EGMPI_RANK"), *s = getenv("EGMPI_SIZE");
    if (j && r && s) {
        snprintf(g_job, sizeof g_job, "%s", j);
        g_rank = atoi(r);
        g_size = atoi(s);
    }
    if (g_size < 1 || g_size > MAXR) die("bad EGMPI_SIZE");
    char name[96];
    snprintf(name, sizeof name, "/egmpi_%s_ctl", g_job);
    size_t total = ctl_bytes() + (size_t)g_size * AR_CHUNK + (size_t)g_size * g_size * MB_SLOTS * MB_BYTES;
    /* every rank opens O_CREAT and truncates to the same size: idempotent, new pages are zero (sparse) */
    char *base = (char *)map_shm(name, total, 1);
    g_ctl = (struct ctl *)base;
    g_ar = base + ctl_bytes();
    g_mbd = g_ar + (size_t)g_size * AR_CHUNK;
    MPI_Barrier(MPI_COMM_WORLD);
    return MPI_SUCCESS;
}

int MPI_Finalize(void) {
    MPI_Barrier(MPI_COMM_WORLD);
    for (int i = 0; i < g_nalloc; i++) shm_unlink(g_alloc[i].name);
    if (!getenv("EGMPI_JOB")) {
        char name[96];
        snprintf(name, sizeof name, "/egmpi_%s_ctl", g_job);
        shm_unlink(name);
    }
    return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = g_size; return MPI_SUCCESS; }

double MPI_Wtime(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Barrier(MPI_Comm c) {
    (void)c;
    if (g_size == 1) return MPI_SUCCESS;
    int gen = atomic_load(&g_ctl->bar_gen);
    if (atomic_fetch_add(&g_ctl->bar_count, 1) == g_size - 1) {
        atomic_store(&g_ctl->bar_count, 0);
        atomic_fetch_add(&g_ctl->bar_gen, 1);
    } else {
        unsigned spins = 0;
        while (atomic_load(&g_ctl->bar_gen) == gen) relax(&spins);
    }
    return MPI_SUCCESS;
}

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr) {
    (void)info;
    if (g_nalloc == MAXALLOC) die("too many MPI_Alloc_mem");
    size_t sz = ((size_t)size + 4095) & ~(size_t)4095;
    if (sz == 0) sz = 4096;
    snprintf(g_alloc[g_nalloc].name, 64, "/egmpi_%s_m%d_%d", g_job, g_rank, g_nalloc);
    void *p = map_shm(g_alloc[g_nalloc].name, sz, 1);
    g_alloc[g_nalloc].ptr = p;
    g_alloc[g_nalloc].size = sz;
    g_nalloc++;
    *(void **)baseptr = p;
    return MPI_SUCCESS;
}

int MPI_Free_mem(void *base) {
    for (int i = 0; i < g_nalloc; i++)
        if (g_alloc[i].ptr == base) { munmap(base, g_alloc[i].size); g_alloc[i].ptr = NULL; }
    return MPI_SUCCESS;
}

int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win) {
    (void)info; (void)comm; (void)size;
    if (g_nwin == MAXWIN) die("too many windows");
    int w = g_nwin++, a = -1;
    for (int i = 0; i < g_nalloc; i++)
        if (g_alloc[i].ptr == base) a = i;
    if (a < 0) die("MPI_Win_create: base must come from MPI_Alloc_mem in this shim");
    struct winrec *rec = &g_ctl->win[w];
    memcpy(rec->name[g_rank], g_alloc[a].name, 64);
    rec->size[g_rank] = (long)g_alloc[a].size;
    rec->disp_unit[g_rank] = disp_unit;
    MPI_Barrier(MPI_COMM_WORLD);
    for (int r = 0; r < g_size; r++) {
        g_win[w].size[r] = (size_t)rec->size[r];
        g_win[w].disp_unit[r] = rec->disp_unit[r];
        g_win[w].peer[r] = (r == g_rank) ? (char *)base : (char *)map_shm(rec->name[r], (size_t)rec->size[r], 0);
    }
    MPI_Barrier(MPI_COMM_WORLD);
    *win = w;
    return MPI_SUCCESS;
}

int MPI_Win_free(MPI_Win *win) { (void)win; MPI_Barrier(MPI_COMM_WORLD); return MPI_SUCCESS; }

/* passive-target shared lock: the reference's readers never lock, so the epoch only has to order the
 * Put before the unlock returns ("remote completion") -- a full fence on a cache-coherent node. */
int MPI_Win_lock(int t, int r, int a, MPI_Win w) { (void)t; (void)r; (void)a; (void)w; return MPI_SUCCESS; }
int MPI_Win_unlock(int r, MPI_Win w) { (void)r; (void)w; atomic_thread_fence(memory_order_seq_cst); return MPI_SUCCESS; }
int MPI_Win_flush(int r, MPI_Win w) { (void)r; (void)w; atomic_thread_fence(memory_order_seq_cst); return MPI_SUCCESS; }

int MPI_Put(const void *origin, int ocount, MPI_Datatype otype, int target, MPI_Aint disp, int tcount,
            MPI_Datatype ttype, MPI_Win w) {
    (void)tcount; (void)ttype;
    size_t nbytes = (size_t)ocount * (size_t)(otype & 0xff);
    size_t off = (size_t)disp * (size_t)g_win[w].disp_unit[target];
    if (off + nbytes > g_win[w].size[target]) die("MPI_Put out of window bounds");
    memcpy(g_win[w].peer[target] + off, origin, nbytes);
    return MPI_SUCCESS;
}

#define SUM_LOOP(T)                                                         \
    do {                                                                    \
        T *out = (T *)((char *)recvbuf + done);                             \
        size_t n = chunk / sizeof(T);                                       \
        for (int r = 0; r < g_size; r++) {                                  \
            const T *in = (const T *)(g_ar + (size_t)r * AR_CHUNK);         \
            if (r == 0) for (size_t i = 0; i < n; i++) out[i] = in[i];      \
            else for (size_t i = 0; i < n; i++) out[i] += in[i];            \
        }                                                                   \
    } while (0)

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
    (void)comm;
    if (op != MPI_SUM) die("MPI_Allreduce: only MPI_SUM");
    size_t esz = (size_t)(type & 0xff), total = (size_t)count * esz, done = 0;
    const char *src = (sendbuf == MPI_IN_PLACE) ? (const char *)recvbuf : (const char *)sendbuf;
    if (g_size == 1) {
        if (src != recvbuf) memcpy(recvbuf, src, total);
        return MPI_SUCCESS;
    }
    while (done < total) {
        size_t chunk = total - done < AR_CHUNK ? total - done : AR_CHUNK;
        memcpy(g_ar + (size_t)g_rank * AR_CHUNK, src + done, chunk);
        MPI_Barrier(MPI_COMM_WORLD);
        /* every rank sums in rank order => bit-identical results everywhere */
        switch (type) {
        case MPI_FLOAT: SUM_LOOP(float); break;
        case MPI_DOUBLE: SUM_LOOP(double); break;
        case MPI_INT: SUM_LOOP(int); break;
        case MPI_LONG: SUM_LOOP(long); break;
        case MPI_SHORT: SUM_LOOP(short); break;
        case MPI_CHAR: SUM_LOOP(signed char); break;
        case MPI_UNSIGNED_CHAR: SUM_LOOP(unsigned char); break;
        default: die("MPI_Allreduce: datatype");
        }
        MPI_Barrier(MPI_COMM_WORLD);
        done += chunk;
    }
    return MPI_SUCCESS;
}

static inline struct mbox *mb_hdr(int dst, int src, int slot) { return &g_ctl->mb[dst][src][slot]; }
static inline char *mb_data(int dst, int src, int slot) {
    return g_mbd + (((size_t)dst * g_size + src) * MB_SLOTS + slot) * MB_BYTES;
}

/* synchronous-mode send: the request completes when the receiver has taken the message */
int MPI_Issend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req) {
    (void)comm;
    size_t nbytes = (size_t)count * (size_t)(type & 0xff);
    if (nbytes > MB_BYTES) die("MPI_Issend: message larger than shim mailbox");
    unsigned spins = 0;
    for (;;) {
        for (int s = 0; s < MB_SLOTS; s++) {
            struct mbox *h = mb_hdr(dest, g_rank, s);
            int expect = 0;
            if (atomic_compare_exchange_strong(&h->state, &expect, 1)) { /* 1 = being filled */
                memcpy(mb_data(dest, g_rank, s), buf, nbytes);
                h->tag = tag;
                h->nbytes = (long)nbytes;
                atomic_store_explicit(&h->state, 2, memory_order_release); /* 2 = ready */
                *req = dest * MB_SLOTS + s + 1;
                return MPI_SUCCESS;
            }
        }
        relax(&spins);
    }
}

int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status *status) {
    (void)comm;
    size_t cap = (size_t)count * (size_t)(type & 0xff);
    unsigned spins = 0;
    for (;;) {
        for (int s = 0; s < MB_SLOTS; s++) {
            struct mbox *h = mb_hdr(g_rank, source, s);
            if (atomic_load_explicit(&h->state, memory_order_acquire) == 2 && h->tag == tag) {
                size_t n = (size_t)h->nbytes < cap ? (size_t)h->nbytes : cap;
                memcpy(buf, mb_data(g_rank, source, s), n);
                if (status) { status->MPI_SOURCE = source; status->MPI_TAG = tag; status->MPI_ERROR = 0; }
                atomic_store_explicit(&h->state, 0, memory_order_release);
                return MPI_SUCCESS;
            }
        }
        relax(&spins);
    }
}

int MPI_Wait(MPI_Request *req, MPI_Status *status) {
    (void)status;
    if (*req <= 0) return MPI_SUCCESS;
    int dest = (*req - 1) / MB_SLOTS, s = (*req - 1) % MB_SLOTS;
    /* NOTE: the slot may be re-used by a later Issend of ours only after we observed state 0 here */
    struct mbox *h = mb_hdr(dest, g_rank, s);
    unsigned spins = 0;
    while (atomic_load_explicit(&h->state, memory_order_acquire) != 0) relax(&spins);
    *req = 0;
    return MPI_SUCCESS;
}
