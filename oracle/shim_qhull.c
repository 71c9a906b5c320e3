/* TEST INFRASTRUCTURE (oracle build shim) — not product code.
 *
 * Backing implementation for oracle/shim/qhull_ra.h: an incremental 3-D convex
 * hull with triangulated faces, enough for the reference mesh compiler's
 * mjCMesh::MakeGraph (reference src/user/user_mesh.cc:1639-1882) to build the
 * hull graph of small meshes (BASELINE config 5: 24-vertex cubelets).  qhull
 * itself is a git-fetched dependency of the reference and is absent here.
 * Vertex/facet enumeration order differs from real qhull; parity is unaffected
 * because the product and the oracle consume the same compiled mjModel. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "qhull_ra.h"

typedef struct { int v[3]; int alive; double n[3]; double d; } Tri;

typedef struct {
  Tri* tri; int ntri, captri;
  vertexT* verts; facetT* facets;
  setT** sets; int nsets;
} Store;

static void cross3(double* r, const double* a, const double* b) {
  r[0] = a[1]*b[2] - a[2]*b[1]; r[1] = a[2]*b[0] - a[0]*b[2]; r[2] = a[0]*b[1] - a[1]*b[0];
}
static double dot3(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }

static void die(qhT* qh) { longjmp(qh->errexit, 1); }

void qh_zero(qhT* qh, FILE* errfile) { (void)errfile; memset(qh, 0, sizeof(*qh)); qh->maxhull = -1; }
void qh_init_A(qhT* qh, FILE* in, FILE* out, FILE* err, int argc, char** argv) {
  (void)qh; (void)in; (void)out; (void)err; (void)argc; (void)argv;
}
void qh_initflags(qhT* qh, char* command) {
  const char* ta = strstr(command, "TA");
  if (ta) qh->maxhull = atoi(ta + 2) + 4;
}
void qh_init_B(qhT* qh, coordT* points, int numpoints, int dim, boolT ismalloc) {
  (void)ismalloc;
  qh->first_point = points; qh->num_points = numpoints; qh->hull_dim = dim;
}

static void add_tri(qhT* qh, Store* s, int a, int b, int c, const double* inside) {
  const double* P = qh->first_point;
  if (s->ntri == s->captri) {
    s->captri = s->captri ? 2*s->captri : 64;
    s->tri = (Tri*)realloc(s->tri, s->captri*sizeof(Tri));
  }
  Tri* t = &s->tri[s->ntri++];
  double e1[3], e2[3];
  for (int k = 0; k < 3; k++) { e1[k] = P[3*b+k] - P[3*a+k]; e2[k] = P[3*c+k] - P[3*a+k]; }
  cross3(t->n, e1, e2);
  double len = sqrt(dot3(t->n, t->n));
  if (len > 0) for (int k = 0; k < 3; k++) t->n[k] /= len;
  t->d = dot3(t->n, P + 3*a);
  t->v[0] = a; t->v[1] = b; t->v[2] = c; t->alive = 1;
  if (dot3(t->n, inside) - t->d > 0) {  /* orient outward */
    t->v[1] = c; t->v[2] = b;
    for (int k = 0; k < 3; k++) t->n[k] = -t->n[k];
    t->d = -t->d;
  }
}

void qh_qhull(qhT* qh) {
  const double* P = qh->first_point;
  int n = qh->num_points;
  if (n < 4 || qh->hull_dim != 3) die(qh);
  Store* s = (Store*)calloc(1, sizeof(Store));
  qh->shim_storage = s;

  /* scale for tolerances */
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) {
    if (P[3*i+k] < lo[k]) lo[k] = P[3*i+k];
    if (P[3*i+k] > hi[k]) hi[k] = P[3*i+k];
  }
  double scale = fmax(hi[0]-lo[0], fmax(hi[1]-lo[1], hi[2]-lo[2]));
  double eps = 1e-10 * (scale > 0 ? scale : 1);

  /* initial simplex: two most distant along the widest axis, then furthest from line, then from plane */
  int ax = 0;
  for (int k = 1; k < 3; k++) if (hi[k]-lo[k] > hi[ax]-lo[ax]) ax = k;
  int i0 = 0, i1 = 0;
  for (int i = 0; i < n; i++) { if (P[3*i+ax] < P[3*i0+ax]) i0 = i; if (P[3*i+ax] > P[3*i1+ax]) i1 = i; }
  if (i0 == i1) die(qh);
  double e[3]; for (int k = 0; k < 3; k++) e[k] = P[3*i1+k] - P[3*i0+k];
  int i2 = -1; double best = eps;
  for (int i = 0; i < n; i++) {
    double w[3], c[3]; for (int k = 0; k < 3; k++) w[k] = P[3*i+k] - P[3*i0+k];
    cross3(c, e, w); double dd = sqrt(dot3(c, c));
    if (dd > best) { best = dd; i2 = i; }
  }
  if (i2 < 0) die(qh);
  double e2[3], nn[3]; for (int k = 0; k < 3; k++) e2[k] = P[3*i2+k] - P[3*i0+k];
  cross3(nn, e, e2); double nl = sqrt(dot3(nn, nn)); for (int k = 0; k < 3; k++) nn[k] /= nl;
  int i3 = -1; best = eps;
  for (int i = 0; i < n; i++) {
    double w[3]; for (int k = 0; k < 3; k++) w[k] = P[3*i+k] - P[3*i0+k];
    double dd = fabs(dot3(nn, w));
    if (dd > best) { best = dd; i3 = i; }
  }
  if (i3 < 0) die(qh);
  double inside[3];
  for (int k = 0; k < 3; k++) inside[k] = 0.25*(P[3*i0+k] + P[3*i1+k] + P[3*i2+k] + P[3*i3+k]);
  add_tri(qh, s, i0, i1, i2, inside); add_tri(qh, s, i0, i1, i3, inside);
  add_tri(qh, s, i0, i2, i3, inside); add_tri(qh, s, i1, i2, i3, inside);

  char* done = (char*)calloc(n, 1);
  done[i0] = done[i1] = done[i2] = done[i3] = 1;
  int nhullv = 4;
  for (;;) {
    /* pick the point furthest outside any live face (Q9-like) */
    int pick = -1; double far = eps;
    for (int i = 0; i < n; i++) {
      if (done[i]) continue;
      double m = -1e300;
      for (int t = 0; t < s->ntri; t++) if (s->tri[t].alive) {
        double dd = dot3(s->tri[t].n, P + 3*i) - s->tri[t].d;
        if (dd > m) m = dd;
      }
      if (m <= eps) { done[i] = 1; continue; }
      if (m > far) { far = m; pick = i; }
    }
    if (pick < 0) break;
    if (qh->maxhull > 0 && nhullv >= qh->maxhull) break;
    done[pick] = 1; nhullv++;
    int nt0 = s->ntri;
    char* vis = (char*)calloc(nt0, 1);
    for (int t = 0; t < nt0; t++) if (s->tri[t].alive)
      vis[t] = (dot3(s->tri[t].n, P + 3*pick) - s->tri[t].d > eps);
    /* horizon: directed edge (a,b) of a visible face whose reverse (b,a) belongs to a non-visible live face */
    for (int t = 0; t < nt0; t++) if (s->tri[t].alive && vis[t]) {
      for (int k = 0; k < 3; k++) {
        int a = s->tri[t].v[k], b = s->tri[t].v[(k+1)%3];
        int horizon = 0;
        for (int u = 0; u < nt0 && !horizon; u++) if (s->tri[u].alive && !vis[u])
          for (int j = 0; j < 3; j++)
            if (s->tri[u].v[j] == b && s->tri[u].v[(j+1)%3] == a) { horizon = 1; break; }
        if (horizon) add_tri(qh, s, a, b, pick, inside);
      }
    }
    for (int t = 0; t < nt0; t++) if (vis[t]) s->tri[t].alive = 0;
    free(vis);
  }
  free(done);
}

void qh_triangulate(qhT* qh) { (void)qh; }

void qh_vertexneighbors(qhT* qh) {
  Store* s = (Store*)qh->shim_storage;
  int n = qh->num_points;
  int nf = 0;
  for (int t = 0; t < s->ntri; t++) if (s->tri[t].alive) nf++;
  int* vid = (int*)malloc(n*sizeof(int));
  for (int i = 0; i < n; i++) vid[i] = -1;
  int nv = 0;
  for (int t = 0; t < s->ntri; t++) if (s->tri[t].alive)
    for (int k = 0; k < 3; k++) if (vid[s->tri[t].v[k]] < 0) vid[s->tri[t].v[k]] = 0;
  for (int i = 0; i < n; i++) if (vid[i] == 0) vid[i] = nv++;
  s->verts = (vertexT*)calloc(nv + 1, sizeof(vertexT));
  s->facets = (facetT*)calloc(nf + 1, sizeof(facetT));
  s->sets = (setT**)calloc(nv + nf, sizeof(setT*));
  s->nsets = 0;
  for (int i = 0; i < n; i++) if (vid[i] >= 0) s->verts[vid[i]].point = qh->first_point + 3*i;
  for (int v = 0; v < nv; v++) { s->verts[v].next = &s->verts[v+1]; s->verts[v+1].previous = &s->verts[v]; }
  int f = 0;
  for (int t = 0; t < s->ntri; t++) if (s->tri[t].alive) {
    setT* vs = (setT*)calloc(1, sizeof(setT) + 4*sizeof(void*));
    vs->maxsize = 3;
    for (int k = 0; k < 3; k++) vs->e[k] = &s->verts[vid[s->tri[t].v[k]]];
    vs->e[3] = NULL;
    s->sets[s->nsets++] = vs;
    s->facets[f].vertices = vs;
    s->facets[f].toporient = 0;   /* v[0..2] already counter-clockwise seen from outside */
    s->facets[f].next = &s->facets[f+1];
    s->facets[f+1].previous = &s->facets[f];
    f++;
  }
  for (int v = 0; v < nv; v++) {
    int cnt = 0;
    for (int g = 0; g < nf; g++) for (int k = 0; k < 3; k++)
      if (s->facets[g].vertices->e[k] == &s->verts[v]) cnt++;
    setT* ns = (setT*)calloc(1, sizeof(setT) + (cnt + 1)*sizeof(void*));
    ns->maxsize = cnt;
    int c = 0;
    for (int g = 0; g < nf; g++) for (int k = 0; k < 3; k++)
      if (s->facets[g].vertices->e[k] == &s->verts[v]) ns->e[c++] = &s->facets[g];
    ns->e[c] = NULL;
    s->sets[s->nsets++] = ns;
    s->verts[v].neighbors = ns;
  }
  free(vid);
  qh->num_vertices = nv; qh->num_facets = nf;
  qh->vertex_list = s->verts; qh->facet_list = s->facets;
}

int qh_pointid(qhT* qh, pointT* point) {
  long off = (long)(point - qh->first_point);
  return (int)(off / 3);
}

void qh_freeqhull(qhT* qh, boolT allmem) {
  (void)allmem;
  Store* s = (Store*)qh->shim_storage;
  if (!s) return;
  for (int i = 0; i < s->nsets; i++) free(s->sets[i]);
  free(s->sets); free(s->verts); free(s->facets); free(s->tri); free(s);
  qh->shim_storage = NULL;
}
void qh_memfreeshort(qhT* qh, int* curlong, int* totlong) { (void)qh; *curlong = 0; *totlong = 0; }
