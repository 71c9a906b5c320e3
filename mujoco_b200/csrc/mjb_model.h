// Host-side model handling for libmjb200 (see mjb_model.cc).
#pragma once
#include <string>
#include <vector>
#include "mjb_types.h"

struct mjModel_;
typedef struct mjModel_ mjModel;

namespace mjb {

struct HostModel {
  DModel dm;                 // pointers refer to ib / db below (host memory)
  std::vector<int> ib;
  std::vector<double> db;
};

void set_error(const std::string& s);
const char* get_error();

mjModel* load_mjb(const char* path);
void free_mjb(mjModel* m);
long model_size(const mjModel* m, const char* name);
int get_option(const mjModel* m, const char* name, double* v);
int set_option(mjModel* m, const char* name, double v);
int check_model(const mjModel* m);
// flatten; nconmax/njmax <= 0 pick defaults.  returns 0 or a negative mjb error code
int build_host_model(const mjModel* m, int nconmax, int njmax, HostModel* out);

}  // namespace mjb
