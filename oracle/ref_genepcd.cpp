// ref_genepcd.cpp — runs the reference's own dataset generator (src/test/genePcd.cpp,
// compiled from where it lies, unmodified) against the shims and hands back the cloud it
// would have written to bridge_ground.pcd.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#include <iostream>
#include <sstream>
#define main gndt_ref_genepcd_main
#include "test/genePcd.cpp"  // via -I/root/reference/src
#undef main

extern "C" size_t gndt_ref_bridge_ground(float *xyzw, size_t cap) {
  std::ostringstream log;
  std::streambuf *old = std::cout.rdbuf(log.rdbuf());
  gndt_ref_genepcd_main(0, nullptr);
  std::cout.rdbuf(old);
  const std::vector<float> &v = pcl::io::captured_cloud();
  size_t n = v.size() / 4 < cap ? v.size() / 4 : cap;
  for (size_t i = 0; i < n * 4; ++i) xyzw[i] = v[i];
  // the generator prints "<i> done": the number of points it assigned
  size_t assigned = 0;
  std::string s = log.str();
  size_t p = s.find(" done");
  if (p != std::string::npos) { size_t b = s.rfind('\n', p); assigned = (size_t)atol(s.c_str() + (b == std::string::npos ? 0 : b + 1)); }
  return assigned;
}
