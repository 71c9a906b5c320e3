/* TEST INFRASTRUCTURE — parity oracle helper, NOT product code.
 *
 * Thin name-based accessors over the reference's own mjModel / mjData, compiled
 * against the reference's public headers and linked to oracle/_ref/libmujoco_ref.so
 * (the UNMODIFIED reference engine).  Python tests drive the reference through
 * ctypes: the mj_* symbols come from libmujoco_ref.so, field lookup from here.
 *
 * Field tables are the reference's own X-macros (include/mujoco/mjxmacro.h:
 * MJMODEL_POINTERS :740, MJDATA_POINTERS :842, MJDATA_ARENA_POINTERS :1023,
 * MJDATA_SCALAR :1032, MJOPTION_FIELDS :23).
 *
 * mjo_rollout mirrors python/mujoco/rollout.cc:67-216 (_unsafe_rollout[_threaded]):
 * per env set FULLPHYSICS state, zero warmstart, loop {set ctrl; mj_step; get state}.
 * It is the CPU baseline ("reference" kind) that bench.py times next to the GPU path.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mujoco/mjxmacro.h>
#include <mujoco/mujoco.h>

#define EXPORT __attribute__((visibility("default")))

static char tcode_mjtNum = 'd';
#define TCODE(type) ( \
  !strcmp(#type, "mjtNum") ? 'd' : !strcmp(#type, "int") ? 'i' : !strcmp(#type, "float") ? 'f' : \
  !strcmp(#type, "mjtByte") ? 'b' : !strcmp(#type, "mjtBool") ? 'b' : !strcmp(#type, "char") ? 'c' : \
  !strcmp(#type, "uintptr_t") ? 'p' : !strcmp(#type, "mjtSize") ? 'q' : !strcmp(#type, "size_t") ? 'q' : \
  !strcmp(#type, "mjContact") ? 'C' : '?')

EXPORT mjModel* mjo_load(const char* path, char* err, int nerr) {
  size_t n = strlen(path);
  (void)tcode_mjtNum;
  if (n > 4 && !strcmp(path + n - 4, ".mjb")) {
    mjModel* m = mj_loadModel(path, NULL);
    if (!m && err) snprintf(err, nerr, "could not load %s", path);
    return m;
  }
  return mj_loadXML(path, NULL, err, nerr);
}

EXPORT long mjo_model_size(const mjModel* m, const char* name) {
#define X(n) if (!strcmp(name, #n)) return (long)m->n;
  MJMODEL_SIZES
#undef X
  return -1;
}

/* returns 0 on success */
EXPORT int mjo_model_field(const mjModel* m, const char* name, void** ptr, char* type, long* nr, long* nc) {
  MJMODEL_POINTERS_PREAMBLE(m)
#define X(T, n, r, c) if (!strcmp(name, #n)) { *ptr = (void*)m->n; *type = TCODE(T); *nr = (long)(m->r); *nc = (long)(c); return 0; }
  MJMODEL_POINTERS
#undef X
  return -1;
}

EXPORT int mjo_data_field(const mjModel* m, const mjData* d, const char* name, void** ptr, char* type, long* nr, long* nc) {
  MJMODEL_POINTERS_PREAMBLE(m)
#define X(T, n, r, c) if (!strcmp(name, #n)) { *ptr = (void*)d->n; *type = TCODE(T); *nr = (long)(m->r); *nc = (long)(c); return 0; }
  MJDATA_POINTERS
#undef X
#undef MJ_M
#undef MJ_D
#define MJ_M(n) m->n
#define MJ_D(n) d->n
#define X(T, n, r, c) if (!strcmp(name, #n)) { *ptr = (void*)d->n; *type = TCODE(T); *nr = (long)(r); *nc = (long)(c); return 0; }
  MJDATA_ARENA_POINTERS
#undef X
#undef MJ_M
#undef MJ_D
#define MJ_M(n) n
#define MJ_D(n) n
#define X(T, n) if (!strcmp(name, #n)) { *ptr = (void*)&d->n; *type = TCODE(T); *nr = 1; *nc = 1; return 0; }
  MJDATA_SCALAR
#undef X
  if (!strcmp(name, "energy")) { *ptr = (void*)d->energy; *type = 'd'; *nr = 2; *nc = 1; return 0; }
  if (!strcmp(name, "solver_niter")) { *ptr = (void*)d->solver_niter; *type = 'i'; *nr = mjNISLAND; *nc = 1; return 0; }
  if (!strcmp(name, "warning_number")) {  /* gathered copy, valid until next call (not thread safe) */
    static int w[mjNWARNING];
    for (int i = 0; i < mjNWARNING; i++) w[i] = d->warning[i].number;
    *ptr = w; *type = 'i'; *nr = mjNWARNING; *nc = 1; return 0;
  }
  return -1;
}

EXPORT int mjo_opt_get(const mjModel* m, const char* name, double* out, int nmax) {
#define X(T, n, sz) if (!strcmp(name, #n)) { out[0] = (double)m->opt.n; return 1; }
#define XVEC(T, n, sz) if (!strcmp(name, #n)) { int k = (sz) < nmax ? (sz) : nmax; for (int i = 0; i < k; i++) out[i] = (double)m->opt.n[i]; return k; }
  MJOPTION_FIELDS
#undef X
#undef XVEC
  return -1;
}

EXPORT int mjo_opt_set(mjModel* m, const char* name, const double* val, int n) {
#define X(T, nm, sz) if (!strcmp(name, #nm)) { m->opt.nm = (T)val[0]; return 1; }
#define XVEC(T, nm, sz) if (!strcmp(name, #nm)) { int k = (sz) < n ? (sz) : n; for (int i = 0; i < k; i++) m->opt.nm[i] = (T)val[i]; return k; }
  MJOPTION_FIELDS
#undef X
#undef XVEC
  return -1;
}

EXPORT int mjo_contact_size(void) { return (int)sizeof(mjContact); }

/* flatten contacts into plain arrays: dist[n], pos[n*3], frame[n*9], geom[n*2], dim[n], efc_address[n],
 * includemargin[n], friction[n*5], solref[n*2], solimp[n*5], exclude[n] */
EXPORT int mjo_contacts(const mjData* d, int nmax, double* dist, double* pos, double* frame, int* geom, int* dim,
                        int* efc_address, double* includemargin, double* friction, double* solref,
                        double* solimp, int* exclude) {
  int n = d->ncon < nmax ? d->ncon : nmax;
  for (int i = 0; i < n; i++) {
    const mjContact* c = d->contact + i;
    dist[i] = c->dist;
    memcpy(pos + 3*i, c->pos, 3*sizeof(double));
    memcpy(frame + 9*i, c->frame, 9*sizeof(double));
    geom[2*i] = c->geom[0]; geom[2*i+1] = c->geom[1];
    dim[i] = c->dim; efc_address[i] = c->efc_address; includemargin[i] = c->includemargin;
    memcpy(friction + 5*i, c->friction, 5*sizeof(double));
    memcpy(solref + 2*i, c->solref, 2*sizeof(double));
    memcpy(solimp + 5*i, c->solimp, 5*sizeof(double));
    exclude[i] = c->exclude;
  }
  return d->ncon;
}

/* ---------------- batched CPU rollout (mirrors python/mujoco/rollout.cc:67-216) ---------------- */
typedef struct {
  const mjModel* m; mjData* d;
  int nbatch, nstep, nstate, nu;
  const double* state0; const double* ctrl; double* state; int* stats; /* per env: [sum ncon, sum nefc, sum niter, nwarn] */
  volatile int* next; int chunk;
  pthread_barrier_t* start;
} RollJob;

static void roll_range(RollJob* j, int lo, int hi) {
  const mjModel* m = j->m; mjData* d = j->d;
  for (int r = lo; r < hi; r++) {
    /* rollout.cc:85-125: defaults for the user inputs the caller does not specify, then the initial state
     * (no mj_resetData per environment - the reference's rollout does not reset either) */
    mju_zero(d->ctrl, m->nu);
    mju_zero(d->qfrc_applied, m->nv);
    mju_zero(d->xfrc_applied, 6*m->nbody);
    for (int i = 0; i < m->neq; i++) d->eq_active[i] = m->eq_active0[i];
    for (int i = 0; i < m->nbody; i++) {
      int id = m->body_mocapid[i];
      if (id >= 0) { mju_copy3(d->mocap_pos + 3*id, m->body_pos + 3*i); mju_copy4(d->mocap_quat + 4*id, m->body_quat + 4*i); }
    }
    mj_setState(m, d, j->state0 + (size_t)r*j->nstate, mjSTATE_FULLPHYSICS);
    mju_zero(d->qacc_warmstart, m->nv);
    for (int i = 0; i < mjNWARNING; i++) d->warning[i].number = 0;
    long sc = 0, se = 0, si = 0;
    for (int t = 0; t < j->nstep; t++) {
      /* rollout.cc:127-155: once any warning fired, pad the remaining outputs with the current state */
      int nwarning = 0;
      for (int i = 0; i < mjNWARNING; i++) if (d->warning[i].number) { nwarning = 1; break; }
      if (nwarning) {
        for (; t < j->nstep; t++)
          if (j->state) mj_getState(m, d, j->state + ((size_t)r*j->nstep + t)*j->nstate, mjSTATE_FULLPHYSICS);
        break;
      }
      if (j->ctrl) mju_copy(d->ctrl, j->ctrl + ((size_t)r*j->nstep + t)*j->nu, j->nu);
      mj_step(m, d);
      sc += d->ncon; se += d->nefc;
      for (int k = 0; k < (d->nisland > 0 ? d->nisland : 1) && k < mjNISLAND; k++) si += d->solver_niter[k];
      if (j->state) mj_getState(m, d, j->state + ((size_t)r*j->nstep + t)*j->nstate, mjSTATE_FULLPHYSICS);
    }
    if (j->stats) {
      int nw = 0; for (int i = 0; i < mjNWARNING; i++) nw += d->warning[i].number;
      j->stats[4*r] = (int)sc; j->stats[4*r+1] = (int)se; j->stats[4*r+2] = (int)si; j->stats[4*r+3] = nw;
    }
  }
}

static void* roll_worker(void* arg) {
  RollJob* j = (RollJob*)arg;
  pthread_barrier_wait(j->start);   /* the clock starts once every worker exists (the reference keeps a persistent ThreadPool) */
  for (;;) {
    int lo = __sync_fetch_and_add(j->next, j->chunk);
    if (lo >= j->nbatch) break;
    int hi = lo + j->chunk < j->nbatch ? lo + j->chunk : j->nbatch;
    roll_range(j, lo, hi);
  }
  return NULL;
}

/* state0 [nbatch][nstate] (mjSTATE_FULLPHYSICS), ctrl [nbatch][nstep][nu] or NULL,
 * state out [nbatch][nstep][nstate] or NULL, stats [nbatch][4] or NULL.  returns seconds elapsed. */
EXPORT double mjo_rollout(const mjModel* m, int nbatch, int nstep, const double* state0, const double* ctrl,
                          double* state, int* stats, int nthread) {
  if (nthread < 1) nthread = 1;
  int nstate = mj_stateSize(m, mjSTATE_FULLPHYSICS);
  volatile int next = 0;
  int chunk = nbatch / (10*nthread); if (chunk < 1) chunk = 1;
  RollJob* jobs = (RollJob*)calloc(nthread, sizeof(RollJob));
  pthread_t* th = (pthread_t*)calloc(nthread, sizeof(pthread_t));
  pthread_barrier_t start;
  pthread_barrier_init(&start, NULL, (unsigned)nthread + 1);
  for (int i = 0; i < nthread; i++) {
    jobs[i] = (RollJob){m, mj_makeData(m), nbatch, nstep, nstate, m->nu, state0, ctrl, state, stats, &next, chunk, &start};
  }
  struct timespec t0, t1;
  for (int i = 0; i < nthread; i++) pthread_create(&th[i], NULL, roll_worker, &jobs[i]);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&start);
  for (int i = 0; i < nthread; i++) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_barrier_destroy(&start);
  for (int i = 0; i < nthread; i++) mj_deleteData(jobs[i].d);
  free(jobs); free(th);
  return (t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec);
}
