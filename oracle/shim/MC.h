// TEST INFRASTRUCTURE (oracle build shim) — not product code.
// Stand-in for MarchingCubeCpp (SDF-plugin meshes only; unused by any BASELINE model).
#ifndef ORACLE_SHIM_MC_H_
#define ORACLE_SHIM_MC_H_
#include <vector>
namespace MC {
typedef double MC_FLOAT;
struct mcVec3f { MC_FLOAT x, y, z; };
struct mcMesh { std::vector<mcVec3f> vertices, normals; std::vector<unsigned int> indices; };
inline void marching_cube(MC_FLOAT*, int, int, int, mcMesh&) {}
}
#endif
