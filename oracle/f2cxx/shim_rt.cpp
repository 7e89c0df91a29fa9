// shim_rt.cpp -- what the translated ISO_C_BINDING shim (fortran/wuming_b200_*.f90 through f2cxx.py) needs beside f90rt.cpp.
// TEST INFRASTRUCTURE ONLY.  `wm_check` of module wuming_b200_c prints the library's message and STOPs; its Fortran text works on
// character pointers (c_f_pointer, write(6,*)), which is outside the translator's subset, so it is provided here with the same
// behaviour: nothing on success, otherwise STOP with "<where>: <wm_last_error()>" (recorded by the entry-point guard).
#include <string>

#include "f90rt.h"

extern "C" const char* wm_last_error(void);      // include/wuming_b200.h -- resolved from whichever library the test loaded

extern "C" void f90rt_wm_check(int* ierr, const char* where) {
  if (*ierr == 0) return;
  const char* msg = wm_last_error();
  throw f90::Stop{std::string(where) + ": " + (msg ? msg : "") + " (code " + std::to_string(*ierr) + ")"};
}
