// oracle/kokkos_stub/Kokkos_Core.hpp — TEST INFRASTRUCTURE ONLY.
//
// A minimal, single-threaded stand-in for the part of the Kokkos API that the reference's HOMMEXX prototypes
// (compute_and_apply_rhs_test/cxx/level_vectorized_ppscan and .../tiled_vectorized_ppscan) use, so that their
// sources can be compiled UNMODIFIED, where they lie under /root/reference, into oracle/_ref/*.so (oracle/Makefile,
// targets `lv` / `tv`) and produce the answers the CUDA path is checked against. Kokkos itself is not in this image.
//
// What is modelled (and only that):
//   View<DataType, Layout, Space, MemoryTraits>  row-major (LayoutRight), runtime extents + strides, managed
//       allocations zero-initialised and reference-counted, unmanaged wraps of raw pointers, rank <= 8;
//       converting construction obeys Kokkos' rule: equal rank, compatible value type, and every extent that is
//       compile-time in BOTH views must be equal (static_assert, like Kokkos' ViewMapping::is_assignable), runtime
//       extents are checked when the assignment happens (std::abort on mismatch, like Kokkos::abort);
//   subview(v, i | ALL, ...), create_mirror_view (same memory space -> the view itself), deep_copy;
//   TeamPolicy<Serial>::member_type with league_rank/team_rank/team_size/team_barrier/team_scratch/thread_scratch;
//   TeamThreadRange / ThreadVectorRange / parallel_for / single(PerThread|PerTeam) as plain serial loops in index
//       order (one thread, one vector lane: the order a Kokkos::Serial build executes them in);
//   the macros KOKKOS_INLINE_FUNCTION, KOKKOS_FORCEINLINE_FUNCTION, KOKKOS_LAMBDA.
// Nothing here restates reference arithmetic.
#ifndef ORACLE_KOKKOS_STUB_CORE_HPP
#define ORACLE_KOKKOS_STUB_CORE_HPP

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_STUB_SERIAL 1

namespace Kokkos {

inline void abort(const char* msg) {
  std::fprintf(stderr, "Kokkos(stub)::abort: %s\n", msg);
  std::abort();
}

// ---- layouts, memory traits, spaces --------------------------------------------------------------------
struct LayoutRight {};
struct LayoutLeft {};

enum MemoryTraitsFlags { Unmanaged = 0x01, RandomAccess = 0x02, Atomic = 0x04, Restrict = 0x08, Aligned = 0x10 };

template <unsigned T>
struct MemoryTraits {
  enum : unsigned { Unmanaged = T & 0x01u, RandomAccess = T & 0x02u, Atomic = T & 0x04u, Restrict = T & 0x08u,
                    Aligned = T & 0x10u };
  using memory_traits = MemoryTraits<T>;
};

struct HostSpace {
  using memory_space = HostSpace;
};

// per-team / per-thread scratch: a bump allocator over a fixed arena shared by every copy of the team member
// (functors take the member by value: the arena must neither move nor be duplicated)
class ScratchMemorySpaceStub {
 public:
  using memory_space = ScratchMemorySpaceStub;
  struct Arena {
    static constexpr size_t kCapacity = size_t(1) << 20;
    std::unique_ptr<char[]> mem;
    size_t used;
    Arena() : mem(new char[kCapacity]()), used(0) {}
  };
  ScratchMemorySpaceStub() {}
  explicit ScratchMemorySpaceStub(const std::shared_ptr<Arena>& arena) : m_arena(arena) {}
  void* get_shmem(size_t bytes) const {
    if (!m_arena) Kokkos::abort("scratch space requested from a team member without scratch");
    const size_t at = (m_arena->used + 63) & ~size_t(63);
    if (at + bytes > Arena::kCapacity) Kokkos::abort("scratch arena exhausted");
    m_arena->used = at + bytes;
    return m_arena->mem.get() + at;
  }

 private:
  std::shared_ptr<Arena> m_arena;
};

struct Serial {
  using execution_space = Serial;
  using memory_space = HostSpace;
  using scratch_memory_space = ScratchMemorySpaceStub;
  using device_type = Serial;
  static int thread_pool_size() { return 1; }
  static int concurrency() { return 1; }
  static void fence() {}
};
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;

inline void initialize() {}
inline void initialize(int&, char**) {}
inline void finalize() {}
inline void fence() {}

// ---- data-type analysis: T*...*[N0][N1]... ---------------------------------------------------------------
namespace Impl {

struct ALL_t {
  constexpr ALL_t() {}
};

template <class T>
struct DataAnalysis {
  using value_type = T;
  static constexpr int rank = 0;
  static constexpr int rank_dynamic = 0;
  static void static_extents(size_t*, int) {}
  static constexpr size_t static_extent(int) { return 0; }
};
template <class T>
struct DataAnalysis<T*> {
  using value_type = typename DataAnalysis<T>::value_type;
  static_assert(DataAnalysis<T>::rank == DataAnalysis<T>::rank_dynamic, "dynamic extents come first");
  static constexpr int rank = DataAnalysis<T>::rank + 1;
  static constexpr int rank_dynamic = DataAnalysis<T>::rank_dynamic + 1;
  static void static_extents(size_t*, int) {}
  static constexpr size_t static_extent(int) { return 0; }
};
template <class T, size_t N>
struct DataAnalysis<T[N]> {
  using value_type = typename DataAnalysis<T>::value_type;
  static constexpr int rank = DataAnalysis<T>::rank + 1;
  static constexpr int rank_dynamic = DataAnalysis<T>::rank_dynamic;
  // T[N] is the OUTERMOST static extent of the remaining type: it sits right after the extents of ... no:
  // `V*[A][B]` is an array of A arrays of B pointers; peeling [A] first, A is the first static extent.
  // static position p (0-based among the static extents) -> extent
  static constexpr int n_static = rank - rank_dynamic;
  static void static_extents(size_t* e, int at) {
    e[at] = N;
    DataAnalysis<T>::static_extents(e, at + 1);
  }
  static constexpr size_t static_extent(int p) { return p == 0 ? N : DataAnalysis<T>::static_extent(p - 1); }
};

// extent of dimension d if compile-time, 0 if run-time
template <class DataType>
constexpr size_t static_extent_of(int d) {
  return d < DataAnalysis<DataType>::rank_dynamic ? 0
                                                  : DataAnalysis<DataType>::static_extent(d - DataAnalysis<DataType>::rank_dynamic);
}

template <class Dst, class Src, int D>
struct DimsAssignable {
  static constexpr bool value =
      (static_extent_of<Dst>(D - 1) == 0 || static_extent_of<Src>(D - 1) == 0 ||
       static_extent_of<Dst>(D - 1) == static_extent_of<Src>(D - 1)) &&
      DimsAssignable<Dst, Src, D - 1>::value;
};
template <class Dst, class Src>
struct DimsAssignable<Dst, Src, 0> {
  static constexpr bool value = true;
};

template <class V, int R>
struct AddPointers {
  using type = typename AddPointers<V, R - 1>::type*;
};
template <class V>
struct AddPointers<V, 0> {
  using type = V;
};

template <class DataType, class Layout, class Space, class Traits>
struct ViewTraitsStub {
  using data_type = DataType;
  using array_layout = Layout;
  using memory_space = Space;
  using memory_traits = Traits;
  using value_type = typename DataAnalysis<DataType>::value_type;
  static constexpr int rank = DataAnalysis<DataType>::rank;
  static constexpr int rank_dynamic = DataAnalysis<DataType>::rank_dynamic;
};

// Kokkos::Impl::ViewMapping<DstTraits, SrcTraits, void>::is_assignable
template <class DstTraits, class SrcTraits, class Specialize = void>
struct ViewMapping {
  using dst_value = typename DstTraits::value_type;
  using src_value = typename SrcTraits::value_type;
  static constexpr bool is_assignable_value =
      std::is_same<dst_value, src_value>::value || std::is_same<dst_value, const src_value>::value;
  static constexpr bool is_assignable_rank = int(DstTraits::rank) == int(SrcTraits::rank);
  static constexpr bool is_assignable_dimension =
      is_assignable_rank &&
      DimsAssignable<typename DstTraits::data_type, typename SrcTraits::data_type,
                     is_assignable_rank ? int(DstTraits::rank) : 0>::value;
  static constexpr bool is_assignable = is_assignable_value && is_assignable_dimension;
};

}  // namespace Impl

constexpr Impl::ALL_t ALL{};

// ---- View ------------------------------------------------------------------------------------------------
template <class DataType, class Layout = LayoutRight, class Space = HostSpace, class Traits = MemoryTraits<0>>
class View {
 public:
  using traits = Impl::ViewTraitsStub<DataType, Layout, Space, Traits>;
  using data_type = DataType;
  using value_type = typename Impl::DataAnalysis<DataType>::value_type;
  using non_const_value_type = typename std::remove_const<value_type>::type;
  using const_value_type = const non_const_value_type;
  using pointer_type = value_type*;
  using reference_type = value_type&;
  using array_layout = Layout;
  using memory_space = Space;
  using memory_traits = Traits;
  using execution_space = Serial;
  using device_type = Serial;
  using size_type = size_t;
  using HostMirror = View<DataType, Layout, HostSpace, MemoryTraits<0>>;
  enum : int { rank = Impl::DataAnalysis<DataType>::rank, rank_dynamic = Impl::DataAnalysis<DataType>::rank_dynamic,
               Rank = rank };
  static_assert(rank <= 8, "stub View: rank <= 8");
  static_assert(std::is_same<Layout, LayoutRight>::value, "stub View: LayoutRight only");

  View() : m_data(nullptr) { clear(); }

  // managed allocation, zero-initialised (Kokkos value-initialises)
  explicit View(const std::string& label, size_t n0 = kUnset, size_t n1 = kUnset, size_t n2 = kUnset, size_t n3 = kUnset,
                size_t n4 = kUnset, size_t n5 = kUnset, size_t n6 = kUnset, size_t n7 = kUnset)
      : m_data(nullptr), m_label(label) {
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
    const size_t n = size();
    non_const_value_type* p = new non_const_value_type[n ? n : 1]();
    m_own = std::shared_ptr<void>(p, [](void* q) { delete[] static_cast<non_const_value_type*>(q); });
    m_data = p;
  }
  explicit View(const char* label, size_t n0 = kUnset, size_t n1 = kUnset, size_t n2 = kUnset, size_t n3 = kUnset,
                size_t n4 = kUnset, size_t n5 = kUnset, size_t n6 = kUnset, size_t n7 = kUnset)
      : View(std::string(label), n0, n1, n2, n3, n4, n5, n6, n7) {}

  // unmanaged wrap of caller memory
  explicit View(pointer_type p, size_t n0 = kUnset, size_t n1 = kUnset, size_t n2 = kUnset, size_t n3 = kUnset,
                size_t n4 = kUnset, size_t n5 = kUnset, size_t n6 = kUnset, size_t n7 = kUnset)
      : m_data(p) {
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
  }

  // a view living in team / thread scratch memory
  explicit View(const ScratchMemorySpaceStub& space, size_t n0 = kUnset, size_t n1 = kUnset, size_t n2 = kUnset,
                size_t n3 = kUnset, size_t n4 = kUnset, size_t n5 = kUnset, size_t n6 = kUnset, size_t n7 = kUnset)
      : m_data(nullptr) {
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
    m_data = static_cast<pointer_type>(space.get_shmem(size() * sizeof(value_type)));
  }

  View(const View&) = default;
  View& operator=(const View&) = default;

  // converting copy: Kokkos' ViewMapping::is_assignable at compile time, runtime extents checked on assignment
  template <class RD, class RL, class RS, class RT,
            class = typename std::enable_if<!std::is_same<View<RD, RL, RS, RT>, View>::value>::type>
  View(const View<RD, RL, RS, RT>& rhs) : m_data(nullptr) {
    using Mapping = Impl::ViewMapping<traits, typename View<RD, RL, RS, RT>::traits, void>;
    static_assert(Mapping::is_assignable, "Incompatible View copy construction");
    assign_from(rhs);
  }
  template <class RD, class RL, class RS, class RT,
            class = typename std::enable_if<!std::is_same<View<RD, RL, RS, RT>, View>::value>::type>
  View& operator=(const View<RD, RL, RS, RT>& rhs) {
    using Mapping = Impl::ViewMapping<traits, typename View<RD, RL, RS, RT>::traits, void>;
    static_assert(Mapping::is_assignable, "Incompatible View copy assignment");
    assign_from(rhs);
    return *this;
  }

  // element access (always returns a reference to value_type; const views of non-const data are writable, as in Kokkos)
  template <class... I>
  reference_type operator()(I... idx) const {
    static_assert(sizeof...(I) == size_t(rank), "stub View: operator() needs one index per dimension");
    const size_t ii[rank ? rank : 1] = {static_cast<size_t>(idx)...};
    size_t off = 0;
    for (int d = 0; d < rank; ++d) {
#ifdef KOKKOS_STUB_BOUNDS_CHECK
      if (ii[d] >= m_ext[d]) {
        std::fprintf(stderr, "stub View '%s': index %zu out of bounds [0,%zu) in dimension %d\n", m_label.c_str(), ii[d],
                     m_ext[d], d);
        std::abort();
      }
#endif
      off += ii[d] * m_str[d];
    }
    return m_data[off];
  }
  reference_type operator[](size_t i) const { return m_data[i * m_str[0]]; }

  pointer_type data() const { return m_data; }
  pointer_type ptr_on_device() const { return m_data; }
  size_t size() const {
    size_t n = 1;
    for (int d = 0; d < rank; ++d) n *= m_ext[d];
    return n;
  }
  size_t span() const { return size(); }
  size_t extent(int d) const { return d < rank ? m_ext[d] : 1; }
  int extent_int(int d) const { return static_cast<int>(extent(d)); }
  size_t dimension_0() const { return extent(0); }
  size_t stride(int d) const { return m_str[d]; }
  const std::string& label() const { return m_label; }
  bool is_allocated() const { return m_data != nullptr; }

  // ---- stub internals (public so that other View instantiations and subview() reach them)
  static constexpr size_t kUnset = ~size_t(0);
  pointer_type m_data;
  size_t m_ext[8], m_str[8];
  std::shared_ptr<void> m_own;
  std::string m_label;

  void clear() {
    for (int d = 0; d < 8; ++d) m_ext[d] = 0, m_str[d] = 0;
    size_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    Impl::DataAnalysis<DataType>::static_extents(st, rank_dynamic);
    for (int d = rank_dynamic; d < rank; ++d) m_ext[d] = st[d];
    contiguous_strides();
  }
  void contiguous_strides() {
    size_t s = 1;
    for (int d = rank - 1; d >= 0; --d) {
      m_str[d] = s;
      s *= m_ext[d];
    }
  }

 private:
  void set_extents(size_t n0, size_t n1, size_t n2, size_t n3, size_t n4, size_t n5, size_t n6, size_t n7) {
    clear();
    const size_t n[8] = {n0, n1, n2, n3, n4, n5, n6, n7};
    for (int d = 0; d < 8; ++d) {
      if (n[d] == kUnset) {
        if (d < rank_dynamic) Kokkos::abort("View constructor: missing run-time extent");
        continue;
      }
      if (d < rank_dynamic) m_ext[d] = n[d];
      else if (d < rank && n[d] != m_ext[d]) Kokkos::abort("View constructor: extent differs from the compile-time one");
      else if (d >= rank && n[d] != 1) Kokkos::abort("View constructor: too many extents");
    }
    contiguous_strides();
  }
  template <class RV>
  void assign_from(const RV& rhs) {
    clear();
    for (int d = 0; d < rank; ++d) {
      if (d >= rank_dynamic && rhs.m_data && m_ext[d] != rhs.m_ext[d])
        Kokkos::abort("View assignment: extent mismatch with a compile-time extent");
      m_ext[d] = rhs.m_ext[d];
      m_str[d] = rhs.m_str[d];
    }
    m_data = rhs.m_data;
    m_own = rhs.m_own;
    m_label = rhs.m_label;
  }
};

// ---- subview ---------------------------------------------------------------------------------------------
namespace Impl {
template <class... A>
struct CountAll;
template <>
struct CountAll<> {
  static constexpr int value = 0;
};
template <class A0, class... A>
struct CountAll<A0, A...> {
  static constexpr int value = (std::is_same<typename std::decay<A0>::type, ALL_t>::value ? 1 : 0) + CountAll<A...>::value;
};
inline bool is_all(const ALL_t&) { return true; }
template <class I>
inline bool is_all(const I&) { return false; }
inline size_t index_of(const ALL_t&) { return 0; }
template <class I>
inline size_t index_of(const I& i) { return static_cast<size_t>(i); }
}  // namespace Impl

template <class D, class L, class S, class T, class... Args>
View<typename Impl::AddPointers<typename View<D, L, S, T>::value_type, Impl::CountAll<Args...>::value>::type, L, S,
     MemoryTraits<Unmanaged>>
subview(const View<D, L, S, T>& v, Args... args) {
  using Src = View<D, L, S, T>;
  static_assert(sizeof...(Args) == size_t(Src::rank), "subview: one argument per dimension");
  using Dst = View<typename Impl::AddPointers<typename Src::value_type, Impl::CountAll<Args...>::value>::type, L, S,
                   MemoryTraits<Unmanaged>>;
  const bool all[] = {Impl::is_all(args)...};
  const size_t idx[] = {Impl::index_of(args)...};
  Dst out;
  size_t off = 0;
  int o = 0;
  for (int d = 0; d < Src::rank; ++d) {
    if (all[d]) {
      out.m_ext[o] = v.m_ext[d];
      out.m_str[o] = v.m_str[d];
      ++o;
    } else {
      if (idx[d] >= v.m_ext[d]) Kokkos::abort("subview: index out of bounds");
      off += idx[d] * v.m_str[d];
    }
  }
  out.m_data = v.m_data + off;
  out.m_own = v.m_own;
  out.m_label = v.m_label;
  return out;
}

// ---- mirrors and copies (one memory space: a mirror is the view itself) ------------------------------------
template <class D, class L, class S, class T>
typename View<D, L, S, T>::HostMirror create_mirror_view(const View<D, L, S, T>& v) {
  typename View<D, L, S, T>::HostMirror m(v);
  return m;
}
template <class D, class L, class S, class T>
typename View<D, L, S, T>::HostMirror create_mirror(const View<D, L, S, T>& v) {
  typename View<D, L, S, T>::HostMirror m(v.label(), v.extent(0), v.rank > 1 ? v.extent(1) : View<D, L, S, T>::kUnset);
  return m;
}

template <class DD, class DL, class DS, class DT, class SD, class SL, class SS, class ST>
void deep_copy(const View<DD, DL, DS, DT>& dst, const View<SD, SL, SS, ST>& src) {
  static_assert(int(View<DD, DL, DS, DT>::rank) == int(View<SD, SL, SS, ST>::rank), "deep_copy: rank mismatch");
  using DV = typename View<DD, DL, DS, DT>::non_const_value_type;
  if (static_cast<const void*>(dst.data()) == static_cast<const void*>(src.data())) return;
  constexpr int R = View<DD, DL, DS, DT>::rank;
  for (int d = 0; d < R; ++d)
    if (dst.extent(d) != src.extent(d)) Kokkos::abort("deep_copy: extent mismatch");
  // generic strided copy in row-major index order
  size_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const size_t n = dst.size();
  for (size_t c = 0; c < n; ++c) {
    size_t od = 0, os = 0;
    for (int d = 0; d < R; ++d) od += idx[d] * dst.m_str[d], os += idx[d] * src.m_str[d];
    const_cast<DV*>(dst.data())[od] = src.data()[os];
    for (int d = R - 1; d >= 0; --d) {
      if (++idx[d] < dst.extent(d)) break;
      idx[d] = 0;
    }
  }
}
template <class DD, class DL, class DS, class DT>
void deep_copy(const View<DD, DL, DS, DT>& dst, const typename View<DD, DL, DS, DT>::non_const_value_type& value) {
  using V = View<DD, DL, DS, DT>;
  constexpr int R = V::rank;
  size_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const size_t n = dst.size();
  for (size_t c = 0; c < n; ++c) {
    size_t od = 0;
    for (int d = 0; d < R; ++d) od += idx[d] * dst.m_str[d];
    const_cast<typename V::non_const_value_type*>(dst.data())[od] = value;
    for (int d = R - 1; d >= 0; --d) {
      if (++idx[d] < dst.extent(d)) break;
      idx[d] = 0;
    }
  }
}

// ---- team policy -------------------------------------------------------------------------------------------
namespace Impl {

class SerialTeamMember {
 public:
  SerialTeamMember(int league_rank, int league_size)
      : m_league_rank(league_rank), m_league_size(league_size),
        m_team_scratch(std::make_shared<ScratchMemorySpaceStub::Arena>()),
        m_thread_scratch(std::make_shared<ScratchMemorySpaceStub::Arena>()) {}
  int league_rank() const { return m_league_rank; }
  int league_size() const { return m_league_size; }
  int team_rank() const { return 0; }
  int team_size() const { return 1; }
  void team_barrier() const {}
  ScratchMemorySpaceStub team_scratch(int) const { return ScratchMemorySpaceStub(m_team_scratch); }
  ScratchMemorySpaceStub thread_scratch(int) const { return ScratchMemorySpaceStub(m_thread_scratch); }
  ScratchMemorySpaceStub team_shmem() const { return team_scratch(0); }

 private:
  int m_league_rank, m_league_size;
  std::shared_ptr<ScratchMemorySpaceStub::Arena> m_team_scratch, m_thread_scratch;
};

template <class ExecSpace>
struct TeamPolicyInternal {
  using member_type = SerialTeamMember;
};

template <class I, class Member>
struct TeamThreadRangeBoundariesStruct {
  I start, end;
  enum : int { increment = 1 };  // Kokkos: the team size (this thread's stride through the range); one thread here
  TeamThreadRangeBoundariesStruct(const Member&, I count) : start(0), end(count) {}
  TeamThreadRangeBoundariesStruct(const Member&, I b, I e) : start(b), end(e) {}
};
template <class I, class Member>
struct ThreadVectorRangeBoundariesStruct {
  I start, end;
  explicit ThreadVectorRangeBoundariesStruct(I count) : start(0), end(count) {}
  ThreadVectorRangeBoundariesStruct(const Member&, I count) : start(0), end(count) {}
  enum : int { increment = 1 };
};
template <class Member>
struct ThreadSingleStruct {
  explicit ThreadSingleStruct(const Member&) {}
};
template <class Member>
struct VectorSingleStruct {
  explicit VectorSingleStruct(const Member&) {}
};

}  // namespace Impl

struct AUTO_t {};
constexpr AUTO_t AUTO{};

struct PerTeamValue {
  int value;
};
struct PerThreadValue {
  int value;
};

template <class ExecSpace = Serial>
class TeamPolicy {
 public:
  using member_type = Impl::SerialTeamMember;
  using execution_space = ExecSpace;
  TeamPolicy(int league_size, int /*team_size*/, int /*vector_length*/ = 1) : m_league_size(league_size) {}
  TeamPolicy(int league_size, const AUTO_t&, int = 1) : m_league_size(league_size) {}
  int league_size() const { return m_league_size; }
  int team_size() const { return 1; }
  TeamPolicy& set_scratch_size(int, const PerTeamValue&) { return *this; }
  TeamPolicy& set_scratch_size(int, const PerThreadValue&) { return *this; }
  TeamPolicy& set_scratch_size(int, const PerTeamValue&, const PerThreadValue&) { return *this; }
  TeamPolicy& set_chunk_size(int) { return *this; }

 private:
  int m_league_size;
};

template <class ExecSpace = Serial>
class RangePolicy {
 public:
  RangePolicy(size_t b, size_t e) : m_begin(b), m_end(e) {}
  size_t begin() const { return m_begin; }
  size_t end() const { return m_end; }

 private:
  size_t m_begin, m_end;
};

inline PerTeamValue PerTeam(int bytes) { return PerTeamValue{bytes}; }
inline PerThreadValue PerThread(int bytes) { return PerThreadValue{bytes}; }
inline Impl::ThreadSingleStruct<Impl::SerialTeamMember> PerTeam(const Impl::SerialTeamMember& m) {
  return Impl::ThreadSingleStruct<Impl::SerialTeamMember>(m);
}
inline Impl::VectorSingleStruct<Impl::SerialTeamMember> PerThread(const Impl::SerialTeamMember& m) {
  return Impl::VectorSingleStruct<Impl::SerialTeamMember>(m);
}

template <class I>
Impl::TeamThreadRangeBoundariesStruct<I, Impl::SerialTeamMember> TeamThreadRange(const Impl::SerialTeamMember& m, I count) {
  return Impl::TeamThreadRangeBoundariesStruct<I, Impl::SerialTeamMember>(m, count);
}
template <class I>
Impl::ThreadVectorRangeBoundariesStruct<I, Impl::SerialTeamMember> ThreadVectorRange(const Impl::SerialTeamMember& m,
                                                                                     I count) {
  return Impl::ThreadVectorRangeBoundariesStruct<I, Impl::SerialTeamMember>(m, count);
}

// nested parallel_for: serial loops in index order
template <class I, class M, class F>
void parallel_for(const Impl::TeamThreadRangeBoundariesStruct<I, M>& r, const F& f) {
  for (I i = r.start; i < r.end; ++i) f(i);
}
template <class I, class M, class F>
void parallel_for(const Impl::ThreadVectorRangeBoundariesStruct<I, M>& r, const F& f) {
  for (I i = r.start; i < r.end; ++i) f(i);
}
template <class M, class F>
void single(const Impl::ThreadSingleStruct<M>&, const F& f) {
  f();
}
template <class M, class F>
void single(const Impl::VectorSingleStruct<M>&, const F& f) {
  f();
}
// top-level dispatch: one team per league rank, in order
template <class E, class F>
void parallel_for(const TeamPolicy<E>& p, const F& f) {
  for (int r = 0; r < p.league_size(); ++r) {
    Impl::SerialTeamMember m(r, p.league_size());
    f(m);
  }
}
template <class E, class F>
void parallel_for(const std::string&, const TeamPolicy<E>& p, const F& f) {
  parallel_for(p, f);
}
template <class E, class F>
void parallel_for(const RangePolicy<E>& p, const F& f) {
  for (size_t i = p.begin(); i < p.end(); ++i) f(static_cast<int>(i));
}
template <class E, class F, class R>
void parallel_reduce(const RangePolicy<E>& p, const F& f, R& result) {
  R acc = R();
  for (size_t i = p.begin(); i < p.end(); ++i) f(static_cast<int>(i), acc);
  result = acc;
}

}  // namespace Kokkos

#endif  // ORACLE_KOKKOS_STUB_CORE_HPP
