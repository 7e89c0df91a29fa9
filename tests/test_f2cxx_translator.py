"""Unit tests of oracle/f2cxx/f2cxx.py, the Fortran 90 -> C++ translator that turns the reference's source files into the
library the oracle is pinned against (tests/test_ref_transpiled.py).  The Fortran below is written for this test (it is not
reference code): one small procedure per language rule the hot path relies on, with the answer the Fortran standard prescribes
computed independently in Python.  A translator that got literal kinds, operator precedence, integer division, loop semantics,
array-section assignment, SAVE, procedure arguments or the rounding-mode scope wrong would fail here before it could mislead the
parity tests."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F2 = os.path.join(ROOT, "oracle", "f2cxx")
sys.path.insert(0, F2)
import f2cxx  # noqa: E402

SRC = r"""
module helper
  implicit none
  private
  public :: helper__scale
  real(8), save :: factor = 3d0
contains
  subroutine helper__scale(a, n, s)
    integer, intent(in)    :: n
    real(8), intent(inout) :: a(0:n-1)
    real(8), intent(in)    :: s
    integer :: i
    do i=0,n-1
      a(i) = a(i)*s*factor
    enddo
  end subroutine helper__scale
end module helper

module t
  use helper
  implicit none
  private
  public :: t__literals, t__loops
  integer, save :: calls = 0
  real(8), parameter :: third = 1d0/3d0
  real(8), allocatable :: w(:)
contains

  subroutine t__literals(out)
    real(8), intent(out) :: out(12)
    integer :: i
    real(8) :: x
    out(1) = 0.1            ! default-real literal: single precision, then promoted
    out(2) = 0.1d0
    out(3) = sqrt(2.)       ! a float32 square root (2d/proj/reconnection/app.f90 has one)
    out(4) = 1/2            ! integer division
    out(5) = -2**2          ! ** binds tighter than unary minus
    out(6) = 2**3**2        ! right associative
    out(7) = 7/2*2.0        ! left to right: (7/2) = 3 in integers, then 6.0
    out(8) = third
    i = 3
    x = 2.5d0
    out(9)  = i/2 + x       ! 1 + 2.5
    out(10) = x**2 + x**3   ! integer powers by multiplication
    out(11) = 1.5e0         ! default-real exponent form = single
    out(12) = 1.0d0-1.0d-17
  end subroutine t__literals

  subroutine t__logic(i, j, res)
    integer, intent(in)  :: i, j
    integer, intent(out) :: res(6)
    res = 0
    if(i==1.and.j==2) res(1) = 1
    if(i.eq.1.or.j.ne.2) res(2) = 1
    if(.not.(i == 1 .and. j == 2)) res(3) = 1
    if(i >= 1 .and. j /= 3) then
      res(4) = 1
    else if(i < 0) then
      res(4) = 2
    else
      res(4) = 3
    endif
    select case(j)
    case(1)
      res(5) = 10
    case(2,3)
      res(5) = 20
    case default
      res(5) = 30
    end select
    res(6) = int(-1.7d0)*100 + int(1.7d0)*10 + floor(-1.2d0) + mod(-7,3)
  end subroutine t__logic

  subroutine t__loops(n, res)
    integer, intent(in)  :: n
    integer, intent(out) :: res(8)
    integer :: i, j, k, cnt
    cnt = 0
    do i=n,1,-2
      cnt = cnt+i
    enddo
    res(1) = cnt
    res(2) = i              ! the do variable after the loop: first value not executed
    do i=5,1
      cnt = -1000           ! zero-trip loop
    enddo
    res(3) = i
    cnt = 0
    outer: do i=1,n
      do j=1,n
        if(j > i) cycle outer
        if(i*j > 12) exit outer
        cnt = cnt+1
      enddo
    enddo outer
    res(4) = cnt
    res(5) = i
    k = 0
    do while(k*k < n)
      k = k+1
    enddo
    res(6) = k
    cnt = 0
    do i=1,10
      if(i == 3) cycle
      if(i == 6) exit
      cnt = cnt+i
    enddo
    res(7) = cnt
    k = n                    ! the bounds of a do loop are evaluated once
    cnt = 0
    do i=1,k
      k = 0
      cnt = cnt+1
    enddo
    res(8) = cnt
  end subroutine t__loops

  subroutine t__sections(a, n, b, m, tot)
    integer, intent(in)    :: n, m
    real(8), intent(inout) :: a(n)
    real(8), intent(inout) :: b(-1:m,2:3)
    real(8), intent(out)   :: tot(4)
    integer :: cnt(0:4)
    a(2:n) = a(1:n-1)               ! the right-hand side is evaluated before any element is stored
    b(-1:m,2) = b(-1:m,3)*2d0
    b(0,:) = -1d0
    cnt(0:4) = 1
    cnt(2) = 5
    tot(1) = sum(a(2:4))
    tot(2) = sum(cnt(3:2))          ! an empty section sums to zero
    tot(3) = sum(cnt(0:4))
    tot(4) = sum(b(1:m,3))+size(a)+max(1,n,3)+min(2.5d0,m*1d0)
  end subroutine t__sections

  subroutine t__saved(res)
    integer, intent(out) :: res(3)
    logical, save :: first = .true.
    integer, save, allocatable :: store(:)
    if(first)then
      allocate(store(-2:2))
      store(-2:2) = 0
      first = .false.
    endif
    calls = calls+1
    store(0) = store(0)+10
    res(1) = calls
    res(2) = store(0)
    res(3) = size(store)
    if(.not.allocated_w()) then
    endif
  end subroutine t__saved

  subroutine t__callback(a, n, op)
    interface
      subroutine op(x, n, s)
        integer, intent(in)    :: n
        real(8), intent(inout) :: x(0:n-1)
        real(8), intent(in)    :: s
      end subroutine op
    end interface
    integer, intent(in)    :: n
    real(8), intent(inout) :: a(n)
    call op(a, n, 2d0)              ! a procedure dummy
    call helper__scale(a(2), n-1, 0.5d0)   ! an element as the start of the actual array; expressions by reference
  end subroutine t__callback

  subroutine t__rounding(a, res)
    use, intrinsic :: ieee_arithmetic
    real(8), intent(in)  :: a(2)
    real(8), intent(out) :: res(2)
    res(1) = a(1)/a(2)
    call ieee_set_rounding_mode(ieee_down)
    res(2) = a(1)/a(2)              ! operands read from memory after the call, like the reference's ieee_down sections
  end subroutine t__rounding

  subroutine t__after_rounding(a, res)
    real(8), intent(in)  :: a(2)
    real(8), intent(out) :: res(1)
    res(1) = a(1)/a(2)              ! the caller's mode (nearest) is back after t__rounding returned
  end subroutine t__after_rounding

  subroutine t__stop(i)
    integer, intent(in) :: i
    if(i > 0)then
      write(6,*)'stopping'
      stop
    endif
  end subroutine t__stop

end module t
"""
# `allocated()` is outside the subset on purpose: the translator must refuse it rather than guess
SRC_OK = SRC.replace("    if(.not.allocated_w()) then\n    endif\n", "")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("f2cxx")
    cpp = d / "t.cpp"
    mods = SRC_OK.split("end module helper\n")
    cpp.write_text(f2cxx.translate([("helper.f90", mods[0] + "end module helper\n"), ("t.f90", mods[1])]))
    so = d / "libt.so"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-frounding-math", "-fPIC", "-shared",
                        "-DF90_BOUNDS", "-I", F2, "-o", str(so), str(cpp), os.path.join(F2, "f90rt.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(str(so))
    L.f90rt_last_stop.restype = C.c_char_p
    return L


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_literal_kinds_precedence_and_integer_arithmetic(lib):
    out = np.zeros(12)
    lib.t__literals(_p(out))
    f32 = np.float32
    assert out[0] == float(f32(0.1)) and out[0] != 0.1
    assert out[1] == 0.1
    assert out[2] == float(np.sqrt(f32(2.0))) and out[2] != np.sqrt(2.0)
    assert out[3] == 0.0
    assert out[4] == -4.0
    assert out[5] == 512.0
    assert out[6] == 6.0
    assert out[7] == 1.0 / 3.0
    assert out[8] == 3.5
    assert out[9] == 2.5 * 2.5 + (2.5 * 2.5) * 2.5
    assert out[10] == 1.5
    assert out[11] == 1.0 - 1.0e-17


def test_logic_select_and_truncation(lib):
    res = np.zeros(6, np.int32)
    for i, j, want in ((1, 2, [1, 1, 0, 1, 20]), (0, 3, [0, 1, 1, 3, 20]), (-1, 7, [0, 1, 1, 2, 30]), (1, 1, [0, 1, 1, 1, 10])):
        lib.t__logic(C.byref(C.c_int(i)), C.byref(C.c_int(j)), _p(res))
        assert list(res[:5]) == want, (i, j, list(res))
    # int() truncates toward zero, floor() rounds down, mod() takes the sign of the dividend
    assert res[5] == (-1) * 100 + 1 * 10 + (-2) + (-1)


def test_do_loop_semantics(lib):
    res = np.zeros(8, np.int32)
    lib.t__loops(C.byref(C.c_int(7)), _p(res))
    assert res[0] == 7 + 5 + 3 + 1 and res[1] == -1          # do i=7,1,-2 leaves i = -1
    assert res[2] == 5                                       # zero-trip loop: the variable holds the start value
    cnt = 0
    for i in range(1, 8):                                    # the named exit / cycle
        stop = False
        for j in range(1, 8):
            if j > i:
                break
            if i * j > 12:
                stop = True
                break
            cnt += 1
        if stop:
            break
    assert res[3] == cnt and res[4] == i
    assert res[5] == 3                                       # smallest k with k*k >= 7
    assert res[6] == 1 + 2 + 4 + 5
    assert res[7] == 7                                       # trip count fixed at loop entry


def test_sections_alias_and_reductions(lib):
    n, m = 6, 3
    a = np.arange(1.0, n + 1)
    b = np.arange(10.0, 10.0 + 2 * (m + 2)).reshape(2, m + 2)       # Fortran b(-1:m, 2:3): first index fastest
    b0 = b.copy()
    tot = np.zeros(4)
    lib.t__sections(_p(a), C.byref(C.c_int(n)), _p(b), C.byref(C.c_int(m)), _p(tot))
    assert list(a) == [1, 1, 2, 3, 4, 5]                            # a shift, not a smear of a(1)
    want = b0.copy()
    want[0, :] = b0[1, :] * 2
    want[:, 1] = -1.0                                               # b(0,:) is the second element of the first axis
    assert np.array_equal(b, want)
    assert tot[0] == 1 + 2 + 3 and tot[1] == 0 and tot[2] == 4 + 5
    assert tot[3] == want[1, 2:2 + m].sum() + n + max(1, n, 3) + min(2.5, float(m))


def test_save_and_first_call_allocation(lib):
    res = np.zeros(3, np.int32)
    for k in (1, 2, 3):
        lib.t__saved(_p(res))
        assert list(res) == [k, 10 * k, 5]


def test_procedure_dummies_and_element_actuals(lib):
    a = np.arange(1.0, 6.0)
    lib.t__callback(_p(a), C.byref(C.c_int(5)), C.cast(lib.helper__scale, C.c_void_p))
    first = np.arange(1.0, 6.0) * 2.0 * 3.0                         # op = helper__scale: x * s * factor
    first[1:] = first[1:] * 0.5 * 3.0                               # second call starts at a(2)
    assert np.array_equal(a, first)


def test_rounding_mode_is_scoped_to_the_procedure(lib):
    res, a = np.zeros(2), np.array([1.0, 10.0])
    lib.t__rounding(_p(a), _p(res))
    assert res[0] == 0.1
    assert res[1] == np.nextafter(0.1, 0.0)                         # the double nearest to 1/10 lies above it: round-down is one ulp below
    after = np.zeros(1)
    lib.t__after_rounding(_p(a), _p(after))
    assert after[0] == 0.1


def test_stop_is_reported_not_fatal(lib):
    n0 = lib.f90rt_stop_count()
    lib.t__stop(C.byref(C.c_int(0)))
    assert lib.f90rt_stop_count() == n0
    lib.t__stop(C.byref(C.c_int(1)))
    assert lib.f90rt_stop_count() == n0 + 1 and b"STOP at t.f90" in lib.f90rt_last_stop()


def test_constructs_outside_the_subset_are_refused():
    with pytest.raises(f2cxx.TranslateError):
        mods = SRC.split("end module helper\n")
        f2cxx.translate([("helper.f90", mods[0] + "end module helper\n"), ("t.f90", mods[1])])
    for bad in ("subroutine s(a)\n real(8) :: a(4)\n a(1:4:2) = 0d0\n end subroutine s",
                "subroutine s(a)\n type(nosuch) :: a\n a%x = 1\n end subroutine s",
                "subroutine s(a)\n real(8), pointer :: a(:)\n end subroutine s",
                "subroutine s(a)\n real(8) :: a(4)\n where(a > 0) a = 0d0\n end subroutine s",
                "subroutine s(a)\n real(8) :: a(4)\n b = 1d0\n end subroutine s",
                "subroutine s(a)\n complex :: a\n end subroutine s"):
        with pytest.raises(f2cxx.TranslateError):
            f2cxx.translate([("bad.f90", "module m\n implicit none\ncontains\n" + bad + "\nend module m\n")])


# ---- the ISO_C_BINDING subset the C-ABI shim under fortran/ is written in (oracle/f2cxx/shim_harness.py) -------------------------
CSRC = r"""
module cbind
  use iso_c_binding
  implicit none
  public
  type, bind(c) :: rec
    integer(c_int)       :: n
    real(c_double)       :: w(3)
    integer(c_long_long) :: big
  end type rec
  type(c_ptr), save :: handle = c_null_ptr
  type(rec), save   :: r
  interface
    function c_fill(p, out) bind(c, name='c_fill_impl') result(ierr)
      import :: c_int, c_ptr, rec
      type(rec), intent(in)    :: p
      type(c_ptr), intent(out) :: out
      integer(c_int)           :: ierr
    end function
    function c_scale(h, a, n, s, tag) bind(c, name='c_scale_impl') result(ierr)
      import :: c_int, c_ptr, c_double, c_long_long
      type(c_ptr), value          :: h, a
      integer(c_int), value       :: n
      real(c_double), value       :: s
      integer(c_long_long), value :: tag
      integer(c_int)              :: ierr
    end function
  end interface
contains
  subroutine cbind__run(a, n, res)
    integer, intent(in)            :: n
    real(8), intent(inout), target :: a(*)
    integer, intent(out)           :: res(4)
    integer(c_int) :: ierr
    res(1) = 0
    if (.not.c_associated(handle)) res(1) = 1
    r%n = n; r%w(1) = 1.5d0; r%w(2) = 2.5d0; r%w(3) = r%w(1) + r%w(2); r%big = 4000000000_8
    ierr = c_fill(r, handle)
    res(2) = ierr
    if (c_associated(handle)) res(1) = res(1) + 10
    ierr = c_scale(handle, c_loc(a), int(n, c_int), real(r%w(3), c_double), int(n, c_long_long) + r%big)
    res(3) = ierr
    ierr = c_scale(handle, c_null_ptr, 0, 0d0, 0_8)
    res(4) = ierr
  end subroutine cbind__run
end module cbind
"""

CIMPL = r"""
#include <cstring>
struct rec_c { int n; double w[3]; long long big; };
static rec_c g_seen;
extern "C" int c_fill_impl(const rec_c* p, void** out) { g_seen = *p; *out = &g_seen; return (int)sizeof(rec_c); }
extern "C" int c_scale_impl(void* h, void* a, int n, double s, long long tag) {
  if (!a) return -7;
  if (h != &g_seen || g_seen.big != 4000000000LL || tag != 4000000000LL + n) return -1;
  for (int i = 0; i < n; ++i) ((double*)a)[i] *= s;
  return g_seen.n * 100 + (int)(g_seen.w[0] + g_seen.w[1] + g_seen.w[2]);
}
"""


def test_iso_c_binding_subset(tmp_path):
    """bind(c) derived types are C structs of the same layout, bind(c) interface functions are C calls with `value` dummies by value
    and the rest by address, c_loc / c_null_ptr / c_associated / int(x, c_int) / real(x, c_double) mean what the standard says"""
    cpp, impl, so = tmp_path / "cbind.cpp", tmp_path / "impl.cpp", tmp_path / "libcbind.so"
    cpp.write_text(f2cxx.translate([("cbind.f90", CSRC)]))
    impl.write_text(CIMPL)
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I", F2, "-o", str(so), str(cpp), str(impl),
                        os.path.join(F2, "f90rt.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(str(so))
    a, res = np.arange(5, dtype=np.float64), np.zeros(4, np.int32)
    L.cbind__run(_p(a), C.byref(C.c_int(5)), _p(res))
    assert list(res) == [11, 40, 5 * 100 + 8, -7]            # unassociated, then associated; sizeof(rec) = 4 (+4) + 24 + 8
    assert np.array_equal(a, np.arange(5) * 4.0)
