/* TEST INFRASTRUCTURE (oracle build shim) — not product code.
 * Stand-in for the reentrant qhull API used by the reference mesh compiler
 * (src/user/user_mesh.cc:1639-1882, mjCMesh::MakeGraph).  Backed by a small
 * exact-enough incremental 3-D convex hull in oracle/shim_qhull.c.  Only mesh
 * geoms need it (BASELINE config 5, cube_3x3x3.xml). */
#ifndef ORACLE_SHIM_QHULL_RA_H_
#define ORACLE_SHIM_QHULL_RA_H_
#include <setjmp.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double coordT;
typedef coordT pointT;
typedef unsigned int boolT;
#define qh_False 0
#define qh_True 1
#define qh_ALL 1
typedef struct setT { int maxsize; void* e[1]; } setT;  /* NULL-terminated pointer array */
typedef struct vertexT vertexT;
typedef struct facetT facetT;
struct vertexT { vertexT* next; vertexT* previous; pointT* point; setT* neighbors; };
struct facetT { facetT* next; facetT* previous; setT* vertices; unsigned toporient; };
typedef struct qhT {
  jmp_buf errexit;
  boolT NOerrexit;
  int num_vertices, num_facets;
  vertexT* vertex_list;  /* terminated by a sentinel whose next==NULL */
  facetT* facet_list;
  pointT* first_point;
  int num_points, hull_dim;
  int maxhull;           /* from "TA<n>" option, -1 if none */
  void* shim_storage;
} qhT;
#define FORALLvertices for (vertex = qh->vertex_list; vertex && vertex->next; vertex = vertex->next)
#define FORALLfacets for (facet = qh->facet_list; facet && facet->next; facet = facet->next)
#define FOREACHsetelement_(type, set, variable) \
  if (((variable = NULL), set)) for (variable##p = (type**)&((set)->e[0]); (variable = *variable##p++);)
void qh_zero(qhT* qh, FILE* errfile);
void qh_init_A(qhT* qh, FILE* in, FILE* out, FILE* err, int argc, char** argv);
void qh_initflags(qhT* qh, char* command);
void qh_init_B(qhT* qh, coordT* points, int numpoints, int dim, boolT ismalloc);
void qh_qhull(qhT* qh);
void qh_triangulate(qhT* qh);
void qh_vertexneighbors(qhT* qh);
int qh_pointid(qhT* qh, pointT* point);
void qh_freeqhull(qhT* qh, boolT allmem);
void qh_memfreeshort(qhT* qh, int* curlong, int* totlong);
#ifdef __cplusplus
}
#endif
#endif
