// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
//
// verilated_cuda_types.h -- data types of the batched (one instance per CUDA
// thread) Verilated model runtime.
//
// Device-side counterparts of the reference's
//   include/verilatedos.h:470-537     (size / mask / bit-position macros)
//   include/verilated.h:116-129       (CData/SData/IData/QData/EData/WData)
//   include/verilated_types.h:81-126  (VlWide<N>)
//   include/verilated_types.h:149-210 (WDataInP / WDataOutP handles)
//   include/verilated_types.h:1426-1509 (VlUnpacked<T,N>)
//   include/verilated_types.h:1854-1977 (VlNBACommitQueue)
//
// Everything here compiles both as CUDA device code (the product path) and as
// plain host C++ (used ONLY by the CPU-side differential tests, see
// tests/optest/) -- there is no host execution path in the product library.
//
//*************************************************************************
#ifndef VERILATOR_VERILATED_CUDA_TYPES_H_
#define VERILATOR_VERILATED_CUDA_TYPES_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VLCU_HD __host__ __device__ __forceinline__
#define VLCU_D __device__ __forceinline__
#define VLCU_UNROLL _Pragma("unroll")
#else
#define VLCU_HD inline __attribute__((always_inline))
#define VLCU_D inline __attribute__((always_inline))
#define VLCU_UNROLL
#endif

#if defined(__CUDA_ARCH__)
#define VLCU_LIKELY(x) (x)
#define VLCU_UNLIKELY(x) (x)
#else
#define VLCU_LIKELY(x) __builtin_expect(!!(x), 1)
#define VLCU_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif
// The emitter prints VL_LIKELY/VL_UNLIKELY around if-conditions (V3EmitCFunc.h:1300-1327)
#ifndef VL_LIKELY
#define VL_LIKELY(x) VLCU_LIKELY(x)
#define VL_UNLIKELY(x) VLCU_UNLIKELY(x)
#endif

//=========================================================================
// Basic integer types (same C storage classes as the host model)

// Narrow values in MEMORY (the SoA block, arrays, the host interface) keep the host model's
// C storage classes; narrow values in the per-thread model image and in temporaries are held in
// full 32-bit registers.  A GPU register is 32 bits wide: with uint8_t/uint16_t scalars every read
// costs a zero-extension (PRMT/LOP3) because the compiler cannot assume the upper register bits
// are clear.  This is safe because V3Clean reasons about cleanliness with a C++ width of at
// least 32 bits (cppWidth(), src/V3Clean.cpp:53-61): emitted code never relies on an 8/16-bit
// container truncating a value, it masks explicitly wherever upper bits matter.
typedef uint8_t VlcuMem8;
typedef uint16_t VlcuMem16;
#ifndef VLCU_NARROW_STORAGE_TYPES
typedef uint32_t CData;  ///< Data representing 'bit' of 1-8 packed bits (32-bit register)
typedef uint32_t SData;  ///< Data representing 'bit' of 9-16 packed bits (32-bit register)
#else
typedef VlcuMem8 CData;
typedef VlcuMem16 SData;
#endif
typedef uint32_t IData;  ///< Data representing 'bit' of 17-32 packed bits
typedef uint64_t QData;  ///< Data representing 'bit' of 33-64 packed bits
typedef uint32_t EData;  ///< Data representing one element of WData array
typedef EData WData;  ///< Data representing >64 packed bits (used as pointer)

// On the device a wide handle is a raw pointer to (thread-private) limbs.
typedef const EData* WDataInP;
typedef EData* WDataOutP;

//=========================================================================
// Size macros (verilatedos.h:470-491)

#define VL_BYTESIZE 8
#define VL_SHORTSIZE 16
#define VL_IDATASIZE 32
#define VL_QUADSIZE 64
#define VL_EDATASIZE 32
#define VL_EDATASIZE_LOG2 5
#define VL_MULS_MAX_WORDS 128  // (verilatedos.h:511)

#define VL_UL(c) (static_cast<IData>(c##UL))
#define VL_ULL(c) (c##ULL)
#define VL_EUL(n) VL_UL(n)

#define VL_BYTES_I(nbits) (((nbits) + (VL_BYTESIZE - 1)) >> 3)
#define VL_WORDS_I(nbits) (((nbits) + (VL_EDATASIZE - 1)) >> VL_EDATASIZE_LOG2)
#define VL_WQ_WORDS_E 2

#define VL_SIZEBITS_I (VL_IDATASIZE - 1)
#define VL_SIZEBITS_Q (VL_QUADSIZE - 1)
#define VL_SIZEBITS_E (VL_EDATASIZE - 1)

// Masks with ones in the low (nbits mod size) positions; a multiple of the
// size (including 0) means "all bits" (verilatedos.h:520-530).
VLCU_HD constexpr IData vlcu_mask_i(int nbits) {
    return (nbits & 31) ? ((1U << (nbits & 31)) - 1U) : ~0U;
}
VLCU_HD constexpr QData vlcu_mask_q(int nbits) {
    return (nbits & 63) ? ((1ULL << (nbits & 63)) - 1ULL) : ~0ULL;
}
#define VL_MASK_I(nbits) vlcu_mask_i(static_cast<int>(nbits))
#define VL_MASK_Q(nbits) vlcu_mask_q(static_cast<int>(nbits))
#define VL_MASK_E(nbits) VL_MASK_I(nbits)

#define VL_BITWORD_I(bit) ((bit) / VL_IDATASIZE)
#define VL_BITWORD_E(bit) ((bit) >> VL_EDATASIZE_LOG2)
#define VL_BITBIT_I(bit) ((bit) & VL_SIZEBITS_I)
#define VL_BITBIT_Q(bit) ((bit) & VL_SIZEBITS_Q)
#define VL_BITBIT_E(bit) ((bit) & VL_SIZEBITS_E)

#define VL_BITISSET_I(data, bit) ((data) & (VL_UL(1) << VL_BITBIT_I(bit)))
#define VL_BITISSET_Q(data, bit) ((data) & (1ULL << VL_BITBIT_Q(bit)))
#define VL_BITISSET_E(data, bit) ((data) & (VL_EUL(1) << VL_BITBIT_E(bit)))
#define VL_BITISSET_W(data, bit) ((data)[VL_BITWORD_E(bit)] & (VL_EUL(1) << VL_BITBIT_E(bit)))
#define VL_BITISSETLIMIT_W(data, width, bit) (((bit) < (width)) && VL_BITISSET_W(data, bit))
#define VL_BITRSHIFT_W(data, bit) ((data)[VL_BITWORD_E(bit)] >> VL_BITBIT_E(bit))

//=========================================================================
// VlWide -- N little-endian 32-bit limbs private to one instance/thread.
// Stored values keep the bits above the declared width at zero; temporaries
// may be dirty there (verilated_types.h:73-78).

template <size_t N_Words>
struct VlWide final {
    static constexpr size_t Words = N_Words;
    EData m_storage[N_Words];

    VLCU_HD EData& operator[](size_t index) { return m_storage[index]; }
    VLCU_HD const EData& operator[](size_t index) const { return m_storage[index]; }
    VLCU_HD EData& at(size_t index) { return m_storage[index]; }
    VLCU_HD const EData& at(size_t index) const { return m_storage[index]; }
    VLCU_HD constexpr size_t size() const { return N_Words; }
    VLCU_HD EData* data() { return &m_storage[0]; }
    VLCU_HD const EData* data() const { return &m_storage[0]; }
    // A VlWide passes directly where a WDataInP/WDataOutP handle is expected
    VLCU_HD operator EData*() { return &m_storage[0]; }
    VLCU_HD operator const EData*() const { return &m_storage[0]; }
    VLCU_HD bool operator==(const VlWide<N_Words>& that) const {
        EData diff = 0;
        VLCU_UNROLL
        for (size_t i = 0; i < N_Words; ++i) diff |= m_storage[i] ^ that.m_storage[i];
        return diff == 0;
    }
    VLCU_HD bool operator!=(const VlWide<N_Words>& that) const { return !(*this == that); }
};

//=========================================================================
// VlUnpacked -- fixed-size unpacked array held in the per-thread model image.
// (Arrays too large for thread-private storage use VlcuGlobalArray of
// verilated_cuda.h instead; the emitter picks per array)

template <typename T_Value, size_t N_Depth>
struct VlUnpacked final {
    T_Value m_storage[N_Depth];

    VLCU_HD T_Value* data() { return &m_storage[0]; }
    VLCU_HD const T_Value* data() const { return &m_storage[0]; }
    VLCU_HD constexpr size_t size() const { return N_Depth; }
    VLCU_HD T_Value& operator[](size_t index) { return m_storage[index]; }
    VLCU_HD const T_Value& operator[](size_t index) const { return m_storage[index]; }
    VLCU_HD void fill(const T_Value& value) {
        for (size_t i = 0; i < N_Depth; ++i) m_storage[i] = value;
    }
    // Change detection helpers used by trigger computation (V3SenExprBuilder.h:253-260)
    VLCU_HD bool neq(const VlUnpacked<T_Value, N_Depth>& that) const {
        bool differs = false;
        for (size_t i = 0; i < N_Depth; ++i) differs |= (m_storage[i] != that.m_storage[i]);
        return differs;
    }
    VLCU_HD void assign(const VlUnpacked<T_Value, N_Depth>& that) {
        for (size_t i = 0; i < N_Depth; ++i) m_storage[i] = that.m_storage[i];
    }
    VLCU_HD bool operator==(const VlUnpacked<T_Value, N_Depth>& that) const { return !neq(that); }
    VLCU_HD bool operator!=(const VlUnpacked<T_Value, N_Depth>& that) const { return neq(that); }
};

//=========================================================================
// VlNBACommitQueue -- deferred array-element writes (V3Delayed's ValueQueue schemes,
// V3Delayed.cpp:47-110), committed in program order at the end of the NBA region
// (verilated_types.h:1854-1977).  Same template head and enqueue()/commit() calls as the host
// class, so the text EmitCFunc prints compiles against it.  Differences a CUDA thread imposes:
//  - the host class grows a std::vector; here the capacity is a template argument (V3EmitCuda
//    derives it from the size of the target array) and an overflow is recorded, raised by the
//    model's eval wrapper as a per-instance fatal (VLCU_FATAL_NBAQ_OVERFLOW), never dropped
//    silently;
//  - the entries live in thread-local memory OUTSIDE the model image (T_Design::Locals, bound
//    like VlcuLocalArray): an indexed array inside the image would push the whole image out of
//    registers;
//  - commit() takes whichever array class the emitter chose for the target (VlUnpacked,
//    VlcuSharedArray, VlcuLocalArray, VlcuGlobalArray, VlcuStateArray, VlcuMdArray): anything
//    indexable with N_Rank operator[] calls.

template <typename T_Target, bool Partial, typename T_Element, size_t N_Rank, size_t N_Capacity = 32>
class VlNBACommitQueue final {
public:
    struct Entry final {
        T_Element value;
        T_Element mask;  // (Partial only)
        uint32_t indices[N_Rank];
    };
    struct Store final {
        Entry pending[N_Capacity];
    };

private:
    Store* m_storep = nullptr;
    uint32_t m_size = 0;
    bool m_overflow = false;

    template <size_t I, typename T_Array>
    VLCU_HD static decltype(auto) at(T_Array&& arr, const uint32_t* idx) {
        if constexpr (I + 1 == N_Rank) {
            return (arr[idx[I]]);
        } else {
            return at<I + 1>(arr[idx[I]], idx);
        }
    }
    template <typename T_Ref, typename T_Val>
    VLCU_HD static void blend(T_Ref& ref, const T_Val& v, const T_Val& m) {
        ref = static_cast<T_Ref>((v & m) | (ref & ~m));
    }
    template <size_t N>
    VLCU_HD static void blend(VlWide<N>& ref, const VlWide<N>& v, const VlWide<N>& m) {
        for (size_t i = 0; i < N; ++i) ref[i] = (v[i] & m[i]) | (ref[i] & ~m[i]);
    }
    template <typename... T_Args>
    VLCU_HD Entry* push(T_Args... indices) {
        static_assert(sizeof...(indices) == N_Rank, "one index per unpacked dimension");
        if (m_size >= N_Capacity) {
            m_overflow = true;
            return nullptr;
        }
        Entry* const ep = &m_storep->pending[m_size++];
        const uint32_t tmp[] = {static_cast<uint32_t>(indices)...};
        for (size_t i = 0; i < N_Rank; ++i) ep->indices[i] = tmp[i];
        return ep;
    }

public:
    VLCU_HD void bind(Store* storep) {
        m_storep = storep;
        m_size = 0;
        m_overflow = false;
    }
    VLCU_HD bool overflowed() const { return m_overflow; }
    VLCU_HD uint32_t size() const { return m_size; }

    // Whole-element form (Partial == false): enqueue(value, i0[, i1...])
    // Partial-element form (Partial == true): enqueue(value, mask, i0[, i1...])
    template <typename... T_Args>
    VLCU_HD void enqueue(const T_Element& value, T_Args... rest) {
        if constexpr (Partial) {
            enqueuePartial(value, rest...);
        } else {
            if (Entry* const ep = push(rest...)) ep->value = value;
        }
    }
    template <typename... T_Args>
    VLCU_HD void enqueuePartial(const T_Element& value, const T_Element& mask, T_Args... indices) {
        if (Entry* const ep = push(indices...)) {
            ep->value = value;
            ep->mask = mask;
        }
    }
    template <typename T_Commit>
    VLCU_HD void commit(T_Commit& target) {
        for (uint32_t n = 0; n < m_size; ++n) {
            const Entry& e = m_storep->pending[n];
            if constexpr (Partial) {
                blend(at<0>(target, e.indices), e.value, e.mask);
            } else {
                at<0>(target, e.indices) = e.value;
            }
        }
        m_size = 0;
    }
};

#endif  // guard
