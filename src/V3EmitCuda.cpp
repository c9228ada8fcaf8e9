// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
// DESCRIPTION: Verilator: Emit the batched CUDA (sm_100a) model of the design
//
// V3EmitCuda prints, from the scheduled AST that V3EmitCFunc/V3EmitCImp print as
// Vtop___024root*.cpp, one translation unit <prefix>__cuda.cu in which
//
//   * struct <prefix>___024root is the per-instance model image (the class
//     EmitCHeader::emitDesignVarDecls prints, V3EmitCHeaders.cpp:73-143), held by ONE CUDA
//     thread: scalars and VlWide values in registers, run-time indexed arrays behind
//     VlcuSharedArray / VlcuLocalArray / VlcuGlobalArray handles;
//   * struct <prefix>__cuda is the design policy include/verilated_cuda_launch.h drives:
//     the structure-of-arrays slot table (OFF_*), load/store between the SoA block in HBM
//     and the image, port row packing, every CFunc reachable from eval() as a
//     `__device__ static` function, and the host tables (members, coverage points,
//     $display formats, $test$plusargs strings, trace declarations);
//   * the statements and expressions inside those functions are printed by the very visitor
//     that prints the C++ model (this class derives from EmitCFunc, V3EmitCFunc.h:118), so
//     every VL_* operator call, cast, mask and temporary is spelled exactly as in
//     Vtop___024root__0.cpp; only what touches the host runtime is redirected:
//        function headers / calls     gain the per-instance context  `VlcuContext& vlCtx`
//        $display/$write              -> vlCtx.writef(format id, args): formatted on the host
//        $test$plusargs               -> vlCtx.testPlusArg(index): resolved at vlcu_create()
//        $dumpfile/$dumpvars          -> host side (vlcu_run trace sampling + vlcu_vcd_write)
//        VL_FINISH_MT/VL_STOP_MT/VL_FATAL_MT, VL_TIME_UNITED_Q, VL_SCOPED_RAND_RESET_*,
//        vlSymsp->_vm_contextp__->...  are macros / a shim of verilated_cuda.h
//
// Two forms of the model image, chosen per design (--cuda-stream / --no-cuda-stream, default: by
// size and hierarchy):
//
//   register form   a single $root class whose scalars fit a thread's registers: the image is a
//                   struct of the thread, loaded from / stored to the SoA block once per launch.
//   tiled form      designs with non-inlined module instances (V3Descope.cpp:95-125,
//                   V3EmitCHeaders.cpp:63-72) or with more state than a thread can hold: the image
//                   stays in HBM.  Every module class becomes a struct whose members are
//                   VlcuLanes<T> = T[VLCU_TILE], one slot per instance of a TILE of VLCU_TILE
//                   instances; <prefix>__Syms holds one such struct per scope, exactly as the C++
//                   Vtop__Syms does, and is the image of one tile.  A member access
//                   `vlSelfRef.sig.v[vlLane]` is then base + constant + lane*sizeof(T): one
//                   coalesced, immediate-offset load or store per warp, no address arithmetic, and
//                   ptxas can tell all members apart.  Functions of a module class take the
//                   scope's struct (vlSelf) and the tile (vlSymsp), so code is shared between the
//                   instances of a module the way the C++ model shares it.
//*************************************************************************

#include "V3PchAstNoMT.h"

#include "V3EmitCuda.h"

#include "V3File.h"

#include "V3EmitC.h"
#include "V3EmitCBase.h"
#include "V3EmitCFunc.h"
#include "V3Stats.h"
#include "V3String.h"

#include <fstream>
#include <functional>
#include <map>
#include <set>
#include <sstream>

VL_DEFINE_DEBUG_FUNCTIONS;

namespace {

enum class CudaStorage : uint8_t { REG, SHARED, LOCAL, GLOBAL, INTERLEAVED, COVER, LANES };

struct CudaMember final {
    const AstVar* varp = nullptr;
    string name;  // C++ member name
    string hier;  // Verilog name
    int kind = 3;  // VlcuMemberKind: 0 in, 1 out, 2 inout, 3 state, 4 coverage, 5 memory
    int elemBytes = 4;
    uint32_t words = 1;  // 32-bit words of one (wide) element, else 1
    uint32_t depth = 1;  // array elements (all dimensions), else 1
    std::vector<uint32_t> dims;  // declared unpacked dimensions, outermost first
    bool isArray = false;
    bool isWide = false;
    uint32_t width = 1;  // bits of one element
    int msb = 0, lsb = 0;
    CudaStorage storage = CudaStorage::REG;
    uint32_t offset = 0;
    string elemCType;  // CData SData IData QData (EData for wide limbs)
    string elemDecl;  // C type of one element as the model declares it
    string memType;  // tiled form: type of one instance's value in memory (VlUnpacked<VlcuMem8, 4> ...)
    bool laneMajor = false;  // tiled form: the elements of one instance are contiguous
    uint32_t elems() const { return depth * words; }
    uint32_t bytes() const { return elems() * elemBytes; }
};

// Which module arrays are indexed with something that is not a constant
// ... and which of those with an index that depends on a design signal (a member of the model),
// i.e. that differs between instances; an index made of function-local variables and constants
// (a loop counter) is the same in every lane of a warp.
class CudaIndexVisitor final : public VNVisitorConst {
    std::set<const AstVar*>& m_runtimeIndexed;
    std::set<const AstVar*>& m_dataIndexed;
    std::set<const AstVar*>& m_wholeRef;  // arrays passed to a function as a whole
    void visit(AstCCall* nodep) override {
        // (the trace functions of array data types are printed as templates over the array class,
        // they take whichever handle the array has)
        const bool traceFunc = nodep->funcp()->name().find("trace") != string::npos;
        for (const AstNode* argp = traceFunc ? nullptr : nodep->argsp(); argp; argp = argp->nextp()) {
            if (const AstVarRef* const refp = VN_CAST(argp, VarRef)) {
                if (VN_IS(refp->varp()->dtypep()->skipRefp(), UnpackArrayDType)) m_wholeRef.insert(refp->varp());
            }
        }
        iterateChildrenConst(nodep);
    }
    static bool dependsOnMember(const AstNode* nodep) {
        bool member = false;
        nodep->foreach([&](const AstVarRef* refp) {
            if (!refp->varp()->isFuncLocal()) member = true;
        });
        return member;
    }
    void visit(AstArraySel* nodep) override {
        if (!VN_IS(nodep->bitp(), Const)) {
            const AstNode* fromp = nodep->fromp();
            while (const AstArraySel* const selp = VN_CAST(fromp, ArraySel)) fromp = selp->fromp();
            if (const AstVarRef* const refp = VN_CAST(fromp, VarRef)) {
                m_runtimeIndexed.insert(refp->varp());
                if (dependsOnMember(nodep->bitp())) m_dataIndexed.insert(refp->varp());
            }
        }
        iterateChildrenConst(nodep);
    }
    void visit(AstNode* nodep) override { iterateChildrenConst(nodep); }

public:
    CudaIndexVisitor(const AstNodeModule* modp, std::set<const AstVar*>& out, std::set<const AstVar*>& dataOut,
                     std::set<const AstVar*>& wholeOut)
        : m_runtimeIndexed{out}
        , m_dataIndexed{dataOut}
        , m_wholeRef{wholeOut} {
        iterateConst(const_cast<AstNodeModule*>(modp));
    }
};

// Persistent bytes the non-slow (per eval) functions reference: scalars and wide values once if
// read / once if written, arrays once per select site ("touched memory words only").
// This is the static S_state_rd / S_state_wr the roofline model of bench.py uses.
class CudaTouchVisitor final : public VNVisitorConst {
    std::map<const AstVar*, std::pair<uint32_t, uint32_t>>& m_sites;  // var -> (read, write) sites
    bool m_underSel = false;
    void visit(AstCFunc* nodep) override {
        if (nodep->slow() || nodep->isTrace() || !nodep->ifdef().empty()) return;
        iterateChildrenConst(nodep);
    }
    void visit(AstArraySel* nodep) override {
        {
            VL_RESTORER(m_underSel);
            m_underSel = true;
            iterateConst(nodep->fromp());
        }
        iterateConst(nodep->bitp());
    }
    void visit(AstVarRef* nodep) override {
        std::pair<uint32_t, uint32_t>& site = m_sites[nodep->varp()];
        const bool isArray = VN_IS(nodep->varp()->dtypep()->skipRefp(), UnpackArrayDType);
        if (nodep->access().isReadOrRW()) site.first = (isArray && m_underSel) ? site.first + 1 : std::max(site.first, 1U);
        if (nodep->access().isWriteOrRW()) site.second = (isArray && m_underSel) ? site.second + 1 : std::max(site.second, 1U);
    }
    void visit(AstNode* nodep) override { iterateChildrenConst(nodep); }

public:
    CudaTouchVisitor(const AstNodeModule* modp, std::map<const AstVar*, std::pair<uint32_t, uint32_t>>& out)
        : m_sites{out} {
        iterateConst(const_cast<AstNodeModule*>(modp));
    }
};

// AstNode::sameNode (the per-node half of AstNode::sameTree, V3Ast.h:530) is protected; the code
// sharing pass needs it node by node, because it compares scope references its own way
struct CudaAstNodePeek final : public AstNode {
    using AstNode::sameNode;
};
static bool cudaSameNode(const AstNode* ap, const AstNode* bp) { return (ap->*(&CudaAstNodePeek::sameNode))(bp); }

// Tiled form: a non-root module class and one of its instances (AstScope)
struct CudaClass final {
    const AstNodeModule* modp = nullptr;
    string name;  // Vtop_sub
    std::vector<CudaMember> members;
};
struct CudaScopeInst final {
    const AstScope* scopep = nullptr;
    size_t classIdx = 0;  // into m_subClasses
    string sym;  // member of <prefix>__Syms: TOP__top__DOT__sub
    string pretty;  // what the Syms constructor passes to ctor(): "top.sub"
};
// One row of the host member table
struct CudaFlatMember final {
    CudaMember m;
    string tableName;  // root: the member's name; other scopes: <sym>.<name>
    string offsetExpr;  // C++ expression of the member's byte offset
};

class EmitCudaModel final : public EmitCFunc {
    // STATE
    const AstNodeModule* const m_topModp;
    const string m_rootName;  // Vtop___024root
    const string m_designName;  // Vtop__cuda
    const string m_symsName;  // Vtop__Syms (tiled form)
    bool m_stream = false;  // tiled form: the image stays in HBM
    std::vector<CudaMember> m_members;  // of the root class
    std::vector<CudaClass> m_subClasses;  // tiled form: the other module classes
    std::vector<CudaScopeInst> m_scopeInsts;  // ... and their instances, in Syms order
    std::set<const AstVar*> m_laneVars;  // tiled form: members stored as VlcuLanes<T>
    std::set<const AstVar*> m_laneArrayVars;  // ... as VlcuLaneArray<T, N> ([element][lane])
    std::vector<CudaFlatMember> m_flat;  // the host member table
    std::map<const AstVar*, size_t> m_memberOf;
    std::set<const AstVar*> m_dataIndexed;  // arrays with an index that differs between instances
    std::set<const AstVar*> m_wholeRef;  // arrays some function takes as a whole (VlUnpacked<>& parameter)
    std::vector<std::pair<string, string>> m_nbaQueues;  // VlNBACommitQueue members: name, device type
    std::vector<string> m_messages;
    std::vector<string> m_plusargs;
    int m_coverBins = 0;  // register form: counters of the root class; tiled form: of all scopes together
    std::map<const AstNodeModule*, int> m_classCoverBins;  // tiled form: uint32_t __Vcoverage[n] per class
    int m_coverBinBase = 0;  // cover table: first counter of the scope whose points are being listed
    string m_coverHierPrefix;  // ... and what its vlNamep adds in front of a point's hierarchy
    uint32_t m_stateBytes = 0;
    bool m_coverTableMode = false;  // visiting _configure_coverage: print table rows
    std::vector<string> m_traceRows;

    // METHODS
    static string quoted(const string& s) { return "\"" + V3OutFormatter::quoteNameControls(s) + "\""; }
    const AstCFunc* findFunc(const string& name) const {
        for (const AstNode* nodep = m_topModp->stmtsp(); nodep; nodep = nodep->nextp()) {
            if (const AstCFunc* const funcp = VN_CAST(nodep, CFunc)) {
                if (funcp->name() == name) return funcp;
            }
        }
        return nullptr;
    }
    bool isDeviceFunc(const AstCFunc* funcp) const {
        // The selection of EmitCImp::emitCFuncImp (V3EmitCImp.cpp:410-425), minus what only the
        // host can run
        if (funcp->isTrace()) return false;
        if (funcp->dpiImportPrototype() || funcp->dpiExportDispatcher()) return false;
        if (funcp->isConstructor() || funcp->isDestructor()) return false;
        if (!funcp->ifdef().empty()) return false;  // VL_DEBUG only (_eval_debug_assertions, dumps)
        const string& name = funcp->name();
        if (name.find("_dump_triggers") != string::npos) return false;
        if (name == "_configure_coverage" || name == "__Vconfigure") return false;
        return true;
    }

    // ---- layout ----------------------------------------------------------------
    // Type of one instance's value of a member in memory (tiled form): the C++ model's type with
    // 1/2-byte storage classes for narrow elements
    static string memTypeOf(const AstNodeDType* dtp) {
        dtp = dtp->skipRefp();
        if (const AstUnpackArrayDType* const adtp = VN_CAST(dtp, UnpackArrayDType)) {
            return "VlUnpacked<" + memTypeOf(adtp->subDTypep()) + ", " + cvtToStr(adtp->elementsConst()) + ">";
        }
        if (dtp->isWide()) return "VlWide<" + cvtToStr(dtp->widthWords()) + ">";
        return dtp->width() <= 8 ? "VlcuMem8" : dtp->width() <= 16 ? "VlcuMem16" : dtp->width() <= 32 ? "IData" : "QData";
    }
    void analyzeVar(const AstVar* varp, const std::set<const AstVar*>& runtimeIndexed,
                    std::vector<CudaMember>& out) {
        CudaMember m;
        m.varp = varp;
        m.name = varp->nameProtect();
        m.hier = varp->prettyName();
        if (varp->isPrimaryIO()) {
            m.kind = varp->isInout() ? 2 : varp->isWritable() ? 1 : 0;
        }
        const AstNodeDType* dtp = varp->dtypep()->skipRefp();
        if (const AstNBACommitQueueDType* const qdtp = VN_CAST(dtp, NBACommitQueueDType)) {
            // V3Delayed's ValueQueue scheme: not persistent (empty between evals), so no slot in
            // the state block; entries in thread-local memory, capacity from the target's size
            uint32_t elems = 1;
            for (const AstNodeDType* tp = qdtp->subDTypep()->skipRefp(); VN_IS(tp, UnpackArrayDType);
                 tp = VN_AS(tp, UnpackArrayDType)->subDTypep()->skipRefp()) {
                elems *= VN_AS(tp, UnpackArrayDType)->elementsConst();
            }
            if (m_stream) {
                varp->v3warn(E_UNSUPPORTED, "Unsupported: --cuda-stream: NBA commit queue " << varp->prettyNameQ());
                return;
            }
            const uint32_t capacity = std::min<uint32_t>(std::max<uint32_t>(2 * elems, 16), 256);
            string type = dtp->cType("", false, false);
            const string::size_type close = type.rfind('>');
            UASSERT_OBJ(close != string::npos, varp, "unexpected VlNBACommitQueue type text");
            type.insert(close, ", " + cvtToStr(capacity));
            m_nbaQueues.emplace_back(varp->nameProtect(), type);
            return;
        }
        if (const AstUnpackArrayDType* const adtp = VN_CAST(dtp, UnpackArrayDType)) {
            // VlUnpacked<VlUnpacked<T, D1>, D0>: kept flat, row-major (VlcuMdArray)
            m.isArray = true;
            m.depth = 1;
            for (const AstUnpackArrayDType* ap = adtp; ap; ap = VN_CAST(dtp, UnpackArrayDType)) {
                m.dims.push_back(ap->elementsConst());
                m.depth *= ap->elementsConst();
                dtp = ap->subDTypep()->skipRefp();
            }
        }
        if (!dtp->isIntegralOrPacked()) {
            varp->v3warn(E_UNSUPPORTED,
                         "Unsupported: --cuda: variable of non-integral type " << varp->prettyNameQ());
            return;
        }
        m.width = dtp->width();
        m.elemDecl = dtp->cType("", false, false);
        if (const AstBasicDType* const bdtp = dtp->basicp()) {
            m.msb = bdtp->hi();
            m.lsb = bdtp->lo();
        } else {
            m.msb = m.width - 1;
        }
        if (dtp->isWide()) {
            m.isWide = true;
            m.words = dtp->widthWords();
            m.elemBytes = 4;
            m.elemCType = "EData";
        } else if (m.width <= 8) {
            // in memory a narrow value keeps the host model's storage class; in the image it is a
            // 32-bit register (include/verilated_cuda_types.h)
            m.elemBytes = 1, m.elemCType = "VlcuMem8";
        } else if (m.width <= 16) {
            m.elemBytes = 2, m.elemCType = "VlcuMem16";
        } else if (m.width <= 32) {
            m.elemBytes = 4, m.elemCType = "IData";
        } else {
            m.elemBytes = 8, m.elemCType = "QData";
        }
        if (m_stream) {
            // Tiled form: everything stays in HBM.  Scalars are [lane]; wide values and arrays keep
            // one instance's elements together ([lane][element]: consecutive accesses of ONE
            // instance share sectors, as VlcuGlobalArray does), except large arrays whose run-time
            // indices are the same in all instances (loop counters), which are [element][lane]
            m.memType = memTypeOf(varp->dtypep());
            m.storage = CudaStorage::LANES;
            m.laneMajor = m.isArray || m.isWide;
            if (m.isArray && !m.isWide && m.dims.size() == 1 && m.bytes() > 2048 && !m_dataIndexed.count(varp)
                && V3EmitCuda::layoutSelection()) {
                m.storage = CudaStorage::INTERLEAVED;
                m.laneMajor = false;
                m_laneArrayVars.insert(varp);
            } else {
                m_laneVars.insert(varp);
                if (!m.isArray && !m.isWide) m_laneScalarType[varp] = m.memType;
            }
            out.push_back(m);
            return;
        }
        if (m.isArray) {
            // Where a per-instance array lives (DESIGN.md "run-time indexed arrays"):
            //   registers      small and only ever indexed with constants (trigger vectors)
            //   shared memory  small and run-time indexed: [index][lane] is bank-conflict free
            //   local memory   medium: thread private, L1 resident
            //   HBM            large memories, instance-major so one instance's accesses share sectors
            const bool runtime = runtimeIndexed.count(varp) != 0;
            const uint32_t bytes = m.bytes();
            if (m_wholeRef.count(varp) && bytes <= 256 && m.dims.size() == 1) {
                // handed to functions as VlUnpacked<>& (trigger vectors of several words): stays a
                // plain member of the image; the callees' loops over it have constant trip counts
                m.storage = CudaStorage::REG;
            } else if (!runtime && m.elems() <= 16 && m.dims.size() == 1) {
                m.storage = CudaStorage::REG;
            } else if (runtime && bytes <= 256 && !m.isWide
                       && m_sharedBytes + m.depth * 4U * 64U <= 16384U) {
                // (a shared element is a 32-bit register image of the value, one column per thread
                // of the 64-thread CTA; all shared arrays of a design together stay below 16 KB so
                // that seven CTAs and the port-row stages still fit an SM)
                m_sharedBytes += m.depth * 4U * 64U;
                m.storage = CudaStorage::SHARED;
            } else if (bytes <= 2048) {
                m.storage = CudaStorage::LOCAL;
            } else if (!m.isWide && m.dims.size() == 1 && V3EmitCuda::layoutSelection()
                       && (!m_dataIndexed.count(varp) || V3EmitCuda::interleaveAll())) {
                // large, and every run-time index is the same in all instances (loop counters):
                // leave it in the SoA block, [element][instance], so that a warp's access coalesces
                m.storage = CudaStorage::INTERLEAVED;
                if (m.elemBytes < 4) m.elemDecl = m.elemCType;
            } else {
                m.storage = CudaStorage::GLOBAL;
                m.kind = 5;
                if (!m.isWide && m.elemBytes < 4) m.elemDecl = m.elemCType;  // HBM layout = host layout
            }
        }
        out.push_back(m);
    }
    // The members of one module class, as EmitCHeader::emitDesignVarDecls selects them
    // (V3EmitCHeaders.cpp:133-135)
    void analyzeModule(const AstNodeModule* modp, std::vector<CudaMember>& out) {
        std::set<const AstVar*> runtimeIndexed;
        { const CudaIndexVisitor visitor{modp, runtimeIndexed, m_dataIndexed, m_wholeRef}; }
        for (const AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
            if (const AstVar* const varp = VN_CAST(nodep, Var)) {
                if (varp->isIO() || varp->isSignal() || varp->isClassMember() || varp->isTemp()
                    || varp->isGenVar()) {
                    if (varp->isStatic()) continue;
                    analyzeVar(varp, runtimeIndexed, out);
                }
            } else if (const AstCell* const cellp = VN_CAST(nodep, Cell)) {
                if (!m_stream) {
                    cellp->v3warn(E_UNSUPPORTED, "Unsupported: --no-cuda-stream: module instance "
                                                     << cellp->prettyNameQ()
                                                     << " was not inlined; use --flatten");
                }
            }
        }
    }
    // Register-resident bytes the root class would need in the register form
    uint32_t registerFormBytes() const {
        uint32_t bytes = 0;
        for (const AstNode* nodep = m_topModp->stmtsp(); nodep; nodep = nodep->nextp()) {
            const AstVar* const varp = VN_CAST(nodep, Var);
            if (!varp || varp->isStatic()) continue;
            if (!(varp->isIO() || varp->isSignal() || varp->isClassMember() || varp->isTemp() || varp->isGenVar())) continue;
            const AstNodeDType* const dtp = varp->dtypep()->skipRefp();
            if (VN_IS(dtp, UnpackArrayDType) || !dtp->isIntegralOrPacked()) continue;
            bytes += dtp->isWide() ? 4 * dtp->widthWords() : dtp->width() <= 32 ? 4 : 8;
        }
        return bytes;
    }
    void collectHierarchy() {
        // Every scope by module, as EmitCSyms collects them (V3EmitCSyms.cpp:745-797)
        for (const AstNodeModule* modp = v3Global.rootp()->modulesp(); modp; modp = VN_AS(modp->nextp(), NodeModule)) {
            if (modp == m_topModp || EmitCUtil::isConstPoolMod(modp)) continue;
            std::vector<const AstScope*> scopes;
            for (const AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
                if (const AstScope* const scopep = VN_CAST(nodep, Scope)) scopes.push_back(scopep);
            }
            if (scopes.empty()) continue;
            CudaClass cls;
            cls.modp = modp;
            cls.name = EmitCUtil::prefixNameProtect(modp);
            m_subClasses.push_back(cls);
            for (const AstScope* const scopep : scopes) {
                CudaScopeInst inst;
                inst.scopep = scopep;
                inst.classIdx = m_subClasses.size() - 1;
                inst.sym = protect(scopep->nameDotless());
                inst.pretty = scopep->prettyName();
                m_scopeInsts.push_back(inst);
            }
        }
        // Syms order: every scope followed by the scopes below it, depth first and by name.  All
        // instances of a module class then see their child scopes at the same distance, which is
        // what lets a cell pointer of the C++ class (`vlSelf->cell->sig`, V3Descope.cpp:95-125)
        // become a compile-time offset (VlcuCell)
        std::stable_sort(m_scopeInsts.begin(), m_scopeInsts.end(),
                         [](const CudaScopeInst& a, const CudaScopeInst& b) { return a.sym < b.sym; });
        std::vector<CudaScopeInst> ordered;
        const std::function<void(const AstScope*)> below = [&](const AstScope* parentp) {
            for (const CudaScopeInst& inst : m_scopeInsts) {
                if (inst.scopep->aboveScopep() != parentp) continue;
                ordered.push_back(inst);
                below(inst.scopep);
            }
        };
        below(v3Global.rootp()->topScopep()->scopep());
        UASSERT_OBJ(ordered.size() == m_scopeInsts.size(), m_topModp, "--cuda: scope not reachable from the top scope");
        m_scopeInsts.swap(ordered);
    }
    // The scope a cell of `parentp` (an instance of the class that declares the cell) leads to
    const CudaScopeInst* childScope(const AstScope* parentp, const AstCell* cellp) const {
        for (const CudaScopeInst& inst : m_scopeInsts) {
            if (inst.scopep->aboveScopep() == parentp && inst.scopep->aboveCellp() == cellp) return &inst;
        }
        return nullptr;
    }
    void computeLayout() {
        collectHierarchy();
        // Which form: the tiled one when the design keeps a module hierarchy, or when its scalars
        // would not fit a thread's registers anyway
        m_stream = V3EmitCuda::streamMode() == 1
                   || (V3EmitCuda::streamMode() < 0 && (!m_scopeInsts.empty() || registerFormBytes() > 1024));
        analyzeModule(m_topModp, m_members);
        if (m_stream) {
            for (CudaClass& cls : m_subClasses) analyzeModule(cls.modp, cls.members);
            findSharedCode();
        }
        // Coverage bins of this module: the counters are one more (local memory) array
        const auto binsOf = [](const AstNodeModule* modp) {
            int bins = 0;
            const_cast<AstNodeModule*>(modp)->foreach([&](AstNodeCoverDecl* declp) {
                if (EmitCParentModule::get(declp->dataDeclThisp()) != modp) return;
                bins = std::max(bins, declp->dataDeclThisp()->localBinNum() + declp->size());
            });
            return bins;
        };
        if (m_stream) {
            // every class keeps its own uint32_t __Vcoverage[n] (V3EmitCHeaders.cpp:153-158), every
            // scope its own copy: the shard's counters are those of all scopes, root first
            m_classCoverBins[m_topModp] = binsOf(m_topModp);
            m_coverBins = m_classCoverBins[m_topModp];
            for (const CudaClass& cls : m_subClasses) m_classCoverBins[cls.modp] = binsOf(cls.modp);
            for (const CudaScopeInst& inst : m_scopeInsts) m_coverBins += m_classCoverBins[m_subClasses[inst.classIdx].modp];
        } else {
            m_coverBins = binsOf(m_topModp);
        }
        // Offsets: largest elements first so that every array start is aligned for its type
        // (an offset is multiplied by the padded instance count, a multiple of 128)
        uint32_t offset = 0;
        for (const int bytes : {8, 4, 2, 1}) {
            for (CudaMember& m : m_members) {
                if (m.elemBytes != bytes) continue;
                m.offset = offset;
                offset += m.bytes();
            }
            if (bytes == 4 && m_coverBins && !m_stream) {
                CudaMember cov;
                cov.name = "__Vcoverage";
                cov.kind = 4;
                cov.elemBytes = 4;
                // per instance: one uint32 per bin; per warp: one uint64 per bin for 32 instances, kept
                // in the same block ([tile][bin], bins / 4 bytes of the block per instance)
                cov.depth = V3EmitCuda::coverPerInstance() ? m_coverBins : (m_coverBins + 15) / 16;
                cov.isArray = true;
                cov.width = 32;
                cov.msb = 31;
                cov.storage = CudaStorage::COVER;  // stays in the SoA block, bumped with RED.ADD
                cov.elemCType = "uint32_t";
                cov.elemDecl = "uint32_t";
                cov.offset = offset;
                offset += cov.bytes();
                m_pendingCoverage = cov;
            }
        }
        if (m_coverBins && !m_stream) m_members.push_back(m_pendingCoverage);
        m_stateBytes = (offset + 7U) & ~7U;
        // Touched bytes per eval (static)
        const auto touched = [this](const AstNodeModule* modp, const std::vector<CudaMember>& members,
                                    uint32_t instances) {
            std::map<const AstVar*, std::pair<uint32_t, uint32_t>> sites;
            { const CudaTouchVisitor visitor{modp, sites}; }
            for (const CudaMember& m : members) {
                if (!m.varp) continue;
                const auto it = sites.find(m.varp);
                if (it == sites.end()) continue;
                const uint32_t elemBytes = m.words * m.elemBytes;
                if (m.isArray) {
                    m_touchedRead += instances * std::min(m.bytes(), it->second.first * elemBytes);
                    m_touchedWrite += instances * std::min(m.bytes(), it->second.second * elemBytes);
                } else {
                    m_touchedRead += it->second.first ? instances * elemBytes : 0;
                    m_touchedWrite += it->second.second ? instances * elemBytes : 0;
                }
            }
        };
        touched(m_topModp, m_members, 1);
        for (size_t c = 0; c < m_subClasses.size(); ++c) {
            uint32_t instances = 0;
            for (const CudaScopeInst& inst : m_scopeInsts) instances += inst.classIdx == c;
            // (members of one scope referenced from another scope's functions are counted where
            // they are declared, once per instance)
            touched(m_subClasses[c].modp, m_subClasses[c].members, instances);
        }
        // Member table order: the primary ports first, in source order (= the order of
        // vlcu_port_info() and of the stimulus / observation row words), then everything else
        {
            std::vector<CudaMember> ordered;
            for (const CudaMember* const mp : portsInSourceOrder()) ordered.push_back(*mp);
            for (const CudaMember& m : m_members) {
                if (!(m.varp && m.kind <= 2 && !m.isArray)) ordered.push_back(m);
            }
            m_members.swap(ordered);
        }
        for (size_t i = 0; i < m_members.size(); ++i) {
            if (m_members[i].varp) m_memberOf[m_members[i].varp] = i;
        }
        // The host member table: the root's members, then (tiled form) every other scope's
        for (const CudaMember& m : m_members) {
            m_flat.push_back(CudaFlatMember{
                m, m.name, m_stream ? "VLCU_TILE_OFF(" + m_symsName + ", TOP, " + m.name + ")" : "T::OFF_" + m.name});
        }
        for (const CudaScopeInst& inst : m_scopeInsts) {
            for (const CudaMember& m : m_subClasses[inst.classIdx].members) {
                CudaFlatMember fm{m, inst.sym + "." + m.name,
                                  "VLCU_TILE_OFF(" + m_symsName + ", " + inst.sym + ", " + m.name + ")"};
                fm.m.hier = inst.pretty + "." + m.hier;
                fm.m.kind = m.kind == 5 ? 5 : 3;  // a port of a sub-module is internal state
                m_flat.push_back(fm);
            }
        }
    }
    CudaMember m_pendingCoverage;
    // Tiled form, inside a leaf function: scalar members held in locals (gather at the top, scatter
    // at the end), keyed by the text of the access
    struct CudaCached final {
        AstVarRef* refp = nullptr;  // a representative reference (prints the member's lvalue)
        string local;  // __Vm<n>
        string ctype;  // CData SData IData QData
        string memType;  // VlcuMem8 ...
        bool load = false;  // the old value may be needed (read, or written on some paths only)
        bool store = false;
    };
    std::map<string, CudaCached> m_cached;
    bool m_caching = false;
    std::set<const AstCFunc*> m_prefetchable;  // leaves with a __prefetch companion
    // Tiled form: code shared between the instances of a module class.  V3Order prints one copy of
    // a scope's functions PER INSTANCE whenever they name another scope absolutely
    // (`vlSymsp->TOP__top__DOT__cpu1__DOT__pic.irq`, V3Descope.cpp:95-125), which they nearly always do;
    // in the tile image such a scope lies at a fixed distance from the function's own scope, the
    // same distance for every instance of a replicated block, so the copies are congruent: one
    // is printed, with the reference relative to vlSelf, and called for all instances.
    std::map<const AstCFunc*, const AstCFunc*> m_sharedAs;  // copy -> the one that is printed
    std::set<const AstCFunc*> m_sharedRep;  // printed with scope references relative to vlSelf
    std::map<const AstScope*, size_t> m_scopeIndex;  // position in m_scopeInsts
    std::map<string, size_t> m_symIndex;  // "(&vlSymsp->TOP__x)" -> position in m_scopeInsts
    std::map<std::pair<size_t, long>, std::set<size_t>> m_relUsed;  // (class, distance in Syms order) of
                                                                    // relative references -> scopes relying on it
    const AstScope* m_relSelfp = nullptr;  // scope of the shared function being printed
    std::map<const AstVarRef*, bool> m_refAbsolute;  // ... its references that every copy spells the same
    std::map<const AstVar*, string> m_laneScalarType;  // members a leaf may cache -> memory type
    uint32_t m_sharedBytes = 0;  // shared memory per CTA the SHARED arrays chosen so far take
    uint32_t m_touchedRead = 0, m_touchedWrite = 0;
    uint32_t m_nSaveOrder = 0;
    uint64_t m_saveCheckval = 0;
    string m_saveUnsupported;

    // ---- model struct ------------------------------------------------------------
    static string handleDecl(const CudaMember& m) {
        const string depth = cvtToStr(m.depth);
        string flat;
        switch (m.storage) {
        case CudaStorage::SHARED: flat = "VlcuSharedArray<" + m.elemDecl + ", " + depth + ">"; break;
        case CudaStorage::LOCAL: flat = "VlcuLocalArray<" + m.elemDecl + ", " + depth + ">"; break;
        case CudaStorage::GLOBAL: flat = "VlcuGlobalArray<" + m.elemDecl + ", " + depth + ">"; break;
        case CudaStorage::INTERLEAVED: flat = "VlcuStateArray<" + m.elemDecl + ", " + depth + ">"; break;
        case CudaStorage::COVER:
            return V3EmitCuda::coverPerInstance() ? "VlcuCovArray<" + depth + "> " + m.name
                                                  : "VlcuWarpCovArray<" + depth + " * 16> " + m.name;
        default: return "";
        }
        if (m.dims.size() <= 1) return flat + " " + m.name;
        string dims;
        for (const uint32_t d : m.dims) dims += ", " + cvtToStr(d);
        return "VlcuMdArray<" + flat + dims + "> " + m.name;
    }
    void emitRootStruct() {
        puts("\n// The model image of one instance (Vtop___024root of the C++ model), held by one thread\n");
        puts("struct " + m_rootName + " final {\n");
        for (const CudaMember& m : m_members) {
            if (m.storage == CudaStorage::REG) {
                puts(m.varp->vlArgType(true, false, false) + ";\n");
            } else {
                puts(handleDecl(m) + ";");
                puts(m.storage == CudaStorage::SHARED  ? "  // run-time indexed: shared memory [index][lane]\n"
                     : m.storage == CudaStorage::LOCAL ? "  // thread-local memory\n"
                     : m.storage == CudaStorage::COVER
                         ? (V3EmitCuda::coverPerInstance() ? "  // stays in HBM as cov[bin][instance], RED.ADD\n"
                                                           : "  // stays in HBM as cov[warp][bin], uint64 per warp of instances\n")
                     : m.storage == CudaStorage::INTERLEAVED
                         ? "  // stays in HBM, [element][instance]: instance-uniform indices coalesce\n"
                                                       : "  // stays in HBM, instance-major\n");
            }
        }
        for (const auto& q : m_nbaQueues) {
            puts(q.second + " " + q.first + ";  // entries in thread-local memory (Locals)\n");
        }
        puts("const char* vlNamep;\n");
        puts("};\n");
    }

    // ---- tiled form: the module classes and Syms as tile images ----------------------------
    void emitTileClass(const AstNodeModule* modp, const string& name, const std::vector<CudaMember>& members) {
        // 256-byte aligned: every 4-byte member of every scope starts a 128-byte line
        for (const AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
            if (const AstCell* const cellp = VN_CAST(nodep, Cell)) puts("struct " + name + "__cell__" + cellp->nameProtect() + ";\n");
        }
        puts("struct alignas(256) " + name + " final {\n");
        // CELLS (V3EmitCHeaders.cpp:63-72): where the C++ class holds a pointer, the child scope
        // sits at a fixed distance in the tile
        for (const AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
            if (const AstCell* const cellp = VN_CAST(nodep, Cell)) {
                puts("VlcuCell<" + EmitCUtil::prefixNameProtect(cellp->modp()) + ", " + name + "__cell__"
                     + cellp->nameProtect() + "> " + cellp->nameProtect() + ";\n");
            }
        }
        puts("const char* vlNamep;  // (one per tile, set by the reset kernel)\n");
        if (const int bins = m_classCoverBins[modp]) {
            puts("VlcuTileCov<" + cvtToStr(bins) + "> __Vcoverage;  // one uint64 counter per bin for the tile's instances\n");
        }
        for (const int bytes : {8, 4, 2, 1}) {
            for (const CudaMember& m : members) {
                if (m.elemBytes != bytes) continue;
                if (m.storage == CudaStorage::INTERLEAVED) {
                    puts("VlcuLaneArray<" + m.elemCType + ", " + cvtToStr(m.depth) + "> " + m.name
                         + ";  // [element][lane]: instance-uniform indices coalesce\n");
                } else {
                    puts("VlcuLanes<" + m.memType + "> " + m.name + ";\n");
                }
            }
        }
        puts("};\n");
    }
    void emitTileStructs() {
        puts("\n// The model image of one TILE of VLCU_TILE instances: the module classes of the C++ model\n");
        puts("// (V3EmitCHeaders.cpp:73-143) with every member widened to one slot per instance\n");
        puts("struct " + m_symsName + ";\n");
        for (const CudaClass& cls : m_subClasses) puts("struct " + cls.name + ";\n");
        for (const CudaClass& cls : m_subClasses) emitTileClass(cls.modp, cls.name, cls.members);
        emitTileClass(m_topModp, m_rootName, m_members);
        puts("\n// One tile = what Vtop__Syms holds for one instance (V3EmitCSyms.cpp:923-982)\n");
        puts("struct alignas(256) " + m_symsName + " final {\n");
        puts(m_rootName + " TOP;\n");
        for (const CudaScopeInst& inst : m_scopeInsts) {
            puts(m_subClasses[inst.classIdx].name + " " + inst.sym + ";\n");
        }
        puts("VlcuSymsCtxTiled _vm_contextp__;  // vlSymsp->_vm_contextp__->... in literal C text\n");
        puts("VLCU_D const char* name() const { return \"TOP\"; }\n");
        puts("};\n");
        // Cell distances: taken from the class's first instance, checked for the others
        const auto cellInfos = [this](const AstNodeModule* modp, const string& name,
                                      const std::vector<std::pair<const AstScope*, string>>& insts) {
            for (const AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
                const AstCell* const cellp = VN_CAST(nodep, Cell);
                if (!cellp) continue;
                const string delta0 = [&]() -> string {
                    const CudaScopeInst* const childp = childScope(insts[0].first, cellp);
                    UASSERT_OBJ(childp, cellp, "--cuda: cell without a scope");
                    return "(static_cast<ptrdiff_t>(__builtin_offsetof(" + m_symsName + ", " + childp->sym
                           + ")) - static_cast<ptrdiff_t>(__builtin_offsetof(" + m_symsName + ", " + insts[0].second + ")))";
                }();
                puts("struct " + name + "__cell__" + cellp->nameProtect() + " final {\n");
                puts("static constexpr ptrdiff_t kSelf = __builtin_offsetof(" + name + ", " + cellp->nameProtect() + ");\n");
                puts("static constexpr ptrdiff_t kDelta = " + delta0 + ";\n");
                puts("};\n");
                for (size_t i = 1; i < insts.size(); ++i) {
                    const CudaScopeInst* const childp = childScope(insts[i].first, cellp);
                    UASSERT_OBJ(childp, cellp, "--cuda: cell without a scope");
                    puts("static_assert(static_cast<ptrdiff_t>(__builtin_offsetof(" + m_symsName + ", " + childp->sym
                         + ")) - static_cast<ptrdiff_t>(__builtin_offsetof(" + m_symsName + ", " + insts[i].second + ")) == "
                         + name + "__cell__" + cellp->nameProtect() + "::kDelta, \"cell distance differs between instances\");\n");
                }
            }
        };
        cellInfos(m_topModp, m_rootName, {{v3Global.rootp()->topScopep()->scopep(), "TOP"}});
        for (size_t c = 0; c < m_subClasses.size(); ++c) {
            std::vector<std::pair<const AstScope*, string>> insts;
            for (const CudaScopeInst& inst : m_scopeInsts) {
                if (inst.classIdx == c) insts.emplace_back(inst.scopep, inst.sym);
            }
            cellInfos(m_subClasses[c].modp, m_subClasses[c].name, insts);
        }
        // Distances shared code relies on: taken from one scope, checked for the others that run it
        for (const auto& rel : m_relUsed) {
            const size_t classIdx = rel.first.first;
            const long distance = rel.first.second;
            const string name = relInfoName(m_subClasses[classIdx].name, distance);
            bool first = true;
            for (const size_t i : rel.second) {
                const size_t t = i + distance;
                const string delta = "(static_cast<ptrdiff_t>(__builtin_offsetof(" + m_symsName + ", " + m_scopeInsts[t].sym
                                     + ")) - static_cast<ptrdiff_t>(__builtin_offsetof(" + m_symsName + ", "
                                     + m_scopeInsts[i].sym + ")))";
                if (first) {
                    puts("struct " + name + " final {\nstatic constexpr ptrdiff_t kDelta = " + delta + ";\n};\n");
                    first = false;
                } else {
                    puts("static_assert(" + delta + " == " + name + "::kDelta, \"scope distance differs between instances\");\n");
                }
            }
        }
    }
    // How the glue (port rows, signature) names a member of the root class
    string rootRef(const CudaMember& m) const {
        return m_stream ? "s.symsp->TOP." + m.name + ".v[s.lane]" : "s." + m.name;
    }

    // ---- SoA load / store ----------------------------------------------------------
    void emitLoadStore(bool load) {
        puts(load ? "VLCU_D static void load(Root& s, const VlcuState& st, uint64_t i) {\n"
                  : "VLCU_D static void store(const Root& s, const VlcuState& st, uint64_t i) {\n");
        for (const CudaMember& m : m_members) {
            if (m.storage == CudaStorage::GLOBAL || m.storage == CudaStorage::COVER
                || m.storage == CudaStorage::INTERLEAVED) {
                continue;
            }
            const string off = "OFF_" + m.name;
            if (!m.isArray && !m.isWide) {
                puts(load ? "s." + m.name + " = st.ld<" + m.elemCType + ">(" + off + ", i);\n"
                          : "st.st<" + m.elemCType + ">(" + off + ", i, s." + m.name + ");\n");
            } else if (!m.isArray) {
                puts(load ? "st.ldW(" + off + ", i, s." + m.name + ");\n"
                          : "st.stW(" + off + ", i, s." + m.name + ");\n");
            } else if (!m.isWide) {
                const string arr = "s." + m.name + (m.dims.size() > 1 ? ".flat()" : "");
                puts("for (uint32_t e = 0; e < " + cvtToStr(m.depth) + "; ++e) ");
                puts(load ? arr + "[e] = st.ldE<" + m.elemCType + ">(" + off + ", e, i);\n"
                          : "st.stE<" + m.elemCType + ">(" + off + ", e, i, " + arr + "[e]);\n");
            } else {
                const string arr = "s." + m.name + (m.dims.size() > 1 ? ".flat()" : "");
                const string words = cvtToStr(m.words);
                puts("for (uint32_t e = 0; e < " + cvtToStr(m.depth) + "; ++e) {\n");
                puts("for (uint32_t w = 0; w < " + words + "; ++w) ");
                puts(load ? arr + "[e][w] = st.ldE<EData>(" + off + ", e * " + words + " + w, i);\n"
                          : "st.stE<EData>(" + off + ", e * " + words + " + w, i, " + arr + "[e][w]);\n");
                puts("}\n");
            }
        }
        puts("}\n");
    }

    // ---- port rows --------------------------------------------------------------------
    static uint32_t rowWords(const CudaMember& m) { return m.isWide ? m.words : (m.elemBytes == 8 ? 2 : 1); }
    static string maskOf(int bits) {
        if (bits >= 32) return "0xffffffffU";
        std::ostringstream os;
        os << "0x" << std::hex << ((1ULL << bits) - 1ULL) << "U";
        return os.str();
    }
    // Ports in SOURCE declaration order (V3VariableOrder has sorted the members by size): the
    // order of the stimulus / observation rows and of the output signature
    std::vector<const CudaMember*> portsInSourceOrder() const {
        std::vector<const CudaMember*> ports;
        for (const CudaMember& m : m_members) {
            if (m.varp && m.kind <= 2 && !m.isArray) ports.push_back(&m);
        }
        std::stable_sort(ports.begin(), ports.end(), [](const CudaMember* ap, const CudaMember* bp) {
            const FileLine* const afl = ap->varp->fileline();
            const FileLine* const bfl = bp->varp->fileline();
            if (afl->lineno() != bfl->lineno()) return afl->lineno() < bfl->lineno();
            return afl->firstColumn() < bfl->firstColumn();
        });
        return ports;
    }
    void emitPortRows() {
        uint32_t inWords = 0, outWords = 0, inBytes = 0, outBytes = 0;
        const std::vector<const CudaMember*> ports = portsInSourceOrder();
        // One 32-bit row word per limb; values are masked to the declared width on ingest (the
        // C++ model assumes clean inputs, V3CCtors.cpp:240-280)
        puts("VLCU_D static void load_inputs_row(Root& s, const uint32_t* rowp, uint64_t stride, uint64_t i) {\n");
        for (const CudaMember* const mp : ports) {
            const CudaMember& m = *mp;
            if (m.kind != 0 && m.kind != 2) continue;
            const string w0 = "rowp[" + cvtToStr(inWords) + " * stride + i]";
            if (m.isWide) {
                for (uint32_t w = 0; w < m.words; ++w) {
                    const int bits = (w == m.words - 1) ? (m.width - 32 * w) : 32;
                    puts(rootRef(m) + "[" + cvtToStr(w) + "] = rowp[" + cvtToStr(inWords + w)
                         + " * stride + i] & " + maskOf(bits) + ";\n");
                }
            } else if (m.elemBytes == 8) {
                puts(rootRef(m) + " = (static_cast<QData>(rowp[" + cvtToStr(inWords + 1) + " * stride + i] & "
                     + maskOf(m.width - 32) + ") << 32) | " + w0 + ";\n");
            } else {
                puts(rootRef(m) + " = static_cast<" + m.elemCType + ">(" + w0 + " & " + maskOf(m.width) + ");\n");
            }
            inWords += rowWords(m);
            inBytes += m.bytes();
        }
        puts("}\n");
        string sigBody;
        puts("VLCU_D static void store_outputs_row(const Root& s, uint32_t* rowp, uint64_t stride, uint64_t i) {\n");
        for (const CudaMember* const mp : ports) {
            const CudaMember& m = *mp;
            if (m.kind != 1 && m.kind != 2) continue;
            for (uint32_t w = 0; w < rowWords(m); ++w) {
                string value;
                if (m.isWide) {
                    value = rootRef(m) + "[" + cvtToStr(w) + "]";
                } else if (m.elemBytes == 8) {
                    value = "static_cast<uint32_t>(" + rootRef(m) + (w ? " >> 32)" : ")");
                } else {
                    value = m_stream ? "static_cast<uint32_t>(" + rootRef(m) + ")" : rootRef(m);
                }
                puts("rowp[" + cvtToStr(outWords + w) + " * stride + i] = " + value + ";\n");
                sigBody += "acc = vlcu_sig_word(acc, " + value + ");\n";
            }
            outWords += rowWords(m);
            outBytes += m.bytes();
        }
        puts("}\n");
        puts("VLCU_D static uint64_t outputs_sig(const Root& s, uint64_t sig) {\nuint32_t acc = VLCU_SIG_ACC_INIT;\n"
             + sigBody + "return vlcu_sig_mix(sig, acc);\n}\n");
        m_sizes = "kInWords = " + cvtToStr(inWords) + ", kOutWords = " + cvtToStr(outWords)
                  + ", kInBytes = " + cvtToStr(inBytes) + ", kOutBytes = " + cvtToStr(outBytes);
    }
    string m_sizes;

    // ---- functions -----------------------------------------------------------------------
    string cudaFuncArgs(const AstCFunc* nodep) {
        // cFuncArgs (V3EmitCBase.cpp:97-133) with the context after the self pointer
        string args;
        if (nodep->isLoose() && !nodep->isStatic()) {
            // the class the function belongs to (V3Descope.cpp:95-125: `vlSelf` is the scope's object)
            args += EmitCUtil::prefixNameProtect(EmitCParentModule::get(nodep)) + "* vlSelf, ";
            if (m_stream) args += m_symsName + "* vlSymsp, ";
            args += "VlcuContext& vlCtx";
        }
        if (!nodep->argTypes().empty()) {
            if (!args.empty()) args += ", ";
            args += nodep->argTypes();
        }
        for (const AstNode* stmtp = nodep->argsp(); stmtp; stmtp = stmtp->nextp()) {
            if (const AstVar* const portp = VN_CAST(stmtp, Var)) {
                if (portp->isIO() && !portp->isFuncReturn()) {
                    if (!args.empty()) args += ", ";
                    args += portp->vlArgType(true, false, true);
                }
            }
        }
        return args;
    }

    // ---- constant pool (V3EmitCConstPool.cpp:44-92): wide constants and V3Table tables as
    // device-resident read-only data, under the names the function bodies use
    bool m_constInitMode = false;
    void emitConstPool() {
        const AstConstPool* const poolp = v3Global.rootp()->constPoolp();
        std::vector<const AstVar*> varps;
        for (AstNode* nodep = poolp->modp()->stmtsp(); nodep; nodep = nodep->nextp()) {
            if (const AstVar* const varp = VN_CAST(nodep, Var)) varps.push_back(varp);
        }
        if (varps.empty()) return;
        std::stable_sort(varps.begin(), varps.end(),
                         [](const AstVar* ap, const AstVar* bp) { return ap->name() < bp->name(); });
        puts("\n// Constant pool\n");
        VL_RESTORER(m_constInitMode);
        m_constInitMode = true;
        for (const AstVar* const varp : varps) {
            const string name = EmitCUtil::topClassName() + "__ConstPool__" + varp->nameProtect();
            putns(varp, "static __device__ const " + varp->dtypep()->cType(name, false, false) + " = ");
            iterateConst(varp->valuep());
            puts(";\n");
        }
    }

    // May this function hold members in locals from entry to exit?  Only if nothing else can
    // see or change them meanwhile and every path runs to the end of the body.
    bool leafCanCache(const AstCFunc* funcp) const {
        if (!funcp->isLoose() || funcp->isStatic() || funcp->slow()) return false;
        if (funcp->rtnTypeVoid() != "void") return false;
        bool ok = true;
        funcp->foreach([&](const AstNode* nodep) {
            if (VN_IS(nodep, NodeCCall) || VN_IS(nodep, CReturn) || VN_IS(nodep, JumpGo)) ok = false;
            if (VN_IS(nodep, CReset)) {  // `local = 0` is fine; a member's reset prints its name as text
                const AstNodeAssign* const assignp = VN_CAST(nodep->backp(), NodeAssign);
                const AstVarRef* const refp = assignp ? VN_CAST(assignp->lhsp(), VarRef) : nullptr;
                if (!refp || !refp->varp()->isFuncLocal()) ok = false;
            }
            // (literal C text may name members or leave the function: VL_FATAL_MT)
            if (VN_IS(nodep, CStmt) || VN_IS(nodep, CExpr) || VN_IS(nodep, CStmtUser) || VN_IS(nodep, CExprUser)) ok = false;
            if (const AstStop* const stopp = VN_CAST(nodep, Stop)) {
                if (stopp->isFatal()) ok = false;
            }
        });
        return ok;
    }
    void collectCached(AstCFunc* funcp) {
        std::set<const AstNode*> topStmts;
        for (const AstNode* stmtp = funcp->stmtsp(); stmtp; stmtp = stmtp->nextp()) topStmts.insert(stmtp);
        uint32_t next = 0;
        funcp->foreach([&](AstVarRef* refp) {
            const auto typeIt = m_laneScalarType.find(refp->varp());
            if (typeIt == m_laneScalarType.end()) return;
            CudaCached& c = m_cached[cacheKey(refp)];
            if (!c.refp) {
                c.refp = refp;
                c.local = "__Vm" + cvtToStr(next++);
                c.memType = typeIt->second;
                c.ctype = c.memType == "VlcuMem8" ? "CData" : c.memType == "VlcuMem16" ? "SData" : c.memType;
            }
            if (refp->access().isReadOrRW()) c.load = true;
            if (refp->access().isWriteOrRW()) {
                c.store = true;
                // written on every path and whole only as the target of a top-level assignment
                const AstNodeAssign* const assignp = VN_CAST(refp->backp(), NodeAssign);
                if (!(assignp && assignp->lhsp() == refp && topStmts.count(assignp))) c.load = true;
            }
        });
        // (A member that is only ever assigned keeps its local too: stored where it is assigned, the
        // store would sit between the loads of everything gathered later -- nvcc cannot know they
        // do not alias -- and the leaf would wait for its loads one group at a time: 73 -> 49 M
        // instance-cycles/s on the VeeR-class design.)
        // --cuda-gather-words N (0 = no limit, the default) gathers only the members a big leaf
        // meets first, up to N 32-bit words; the rest are read and written in place
        if (!V3EmitCuda::gatherWords()) return;
        size_t words = 0;
        std::vector<std::pair<unsigned long, string>> order;
        for (const auto& it : m_cached) order.emplace_back(std::stoul(it.second.local.substr(4)), it.first);
        std::sort(order.begin(), order.end());
        for (const auto& entry : order) {
            words += m_cached[entry.second].ctype == "QData" ? 2 : 1;
            if (words > V3EmitCuda::gatherWords()) m_cached.erase(entry.second);
        }
    }

    // ---- sharing between instances ------------------------------------------------------------
    // Position, relative to `selfp`, of the scope an absolute self pointer names; 0 with ok=false
    // if it is TOP / this / a cell (those print the same for every instance anyway)
    long relDistance(const VSelfPointerText& self, const AstScope* selfp, bool& isAbs) const {
        isAbs = false;
        if (!self.isVlSym()) return 0;
        const auto it = m_symIndex.find(self.asString());
        if (it == m_symIndex.end()) return 0;  // (&vlSymsp->TOP)
        isAbs = true;
        return static_cast<long>(it->second) - static_cast<long>(m_scopeIndex.at(selfp));
    }
    // (refAp: the reference in the function that would be printed; how it is to be printed --
    // absolute when every copy names the very same scope, else relative -- must come out the same
    // against every copy)
    bool congruentSelf(const AstVarRef* refAp, const VSelfPointerText& a, const AstScope* scopeAp,
                       const VSelfPointerText& b, const AstScope* scopeBp, std::set<long>& rels,
                       std::map<const AstVarRef*, bool>& modes) const {
        bool absA, absB;
        const long da = relDistance(a, scopeAp, absA), db = relDistance(b, scopeBp, absB);
        if (absA != absB) return false;
        if (!absA) return a == b;
        const bool sameScope = a == b;
        const auto known = m_refAbsolute.find(refAp);
        if (known != m_refAbsolute.end() && known->second != sameScope) return false;
        modes[refAp] = sameScope;
        if (sameScope) return true;
        if (da != db) return false;
        const size_t ia = m_scopeIndex.at(scopeAp) + da, ib = m_scopeIndex.at(scopeBp) + db;
        if (m_scopeInsts[ia].classIdx != m_scopeInsts[ib].classIdx) return false;
        rels.insert(da);
        return true;
    }
    // Are two functions of the same class the same code up to the scope they run in?
    bool congruent(const AstNode* ap, const AstNode* bp, const AstScope* scopeAp, const AstScope* scopeBp,
                   std::map<const AstVar*, const AstVar*>& locals, std::set<long>& rels,
                   std::map<const AstVarRef*, bool>& modes, bool next) const {
        for (; ap || bp; ap = ap->nextp(), bp = bp->nextp()) {
#define return_false(why) do { if (getenv("VLCU_DEBUG_SHARE")) std::cerr << "  mismatch " << why << " at " << (ap ? ap->typeName() : "null") << " " << (ap ? ap->fileline()->ascii() : string{}) << "\n"; return false; } while (0)
            if (!ap || !bp || ap->type() != bp->type()) return_false("type");
            if (const AstVarRef* const ra = VN_CAST(ap, VarRef)) {
                const AstVarRef* const rb = VN_AS(bp, VarRef);
                if (ra->access() != rb->access()) return_false("access");
                if (ra->varp()->isFuncLocal() != rb->varp()->isFuncLocal()) return_false("local-vs-member");
                if (ra->varp()->isFuncLocal()) {
                    const auto it = locals.find(ra->varp());
                    if (it == locals.end() || it->second != rb->varp()) return_false("local var " + ra->varp()->name());
                } else {
                    if (ra->varp() != rb->varp()) return_false("member " + ra->varp()->name() + " vs " + rb->varp()->name());
                    if (!congruentSelf(ra, ra->selfPointer(), scopeAp, rb->selfPointer(), scopeBp, rels, modes)) return_false("scope " + ra->selfPointer().asString() + " vs " + rb->selfPointer().asString());
                }
            } else if (const AstVar* const va = VN_CAST(ap, Var)) {
                const AstVar* const vb = VN_AS(bp, Var);
                if (va->name() != vb->name() || !va->dtypep()->skipRefp()->similarDType(vb->dtypep()->skipRefp())) return_false("var decl");
                locals[va] = vb;
            } else if (const AstNodeCCall* const ca = VN_CAST(ap, NodeCCall)) {
                // (a leaf calls nothing but shared static helpers)
                if (ca->funcp() != VN_AS(bp, NodeCCall)->funcp() || !ca->funcp()->isStatic()) return_false("call");
            } else {
                if (!cudaSameNode(ap, bp)) return_false("sameNode");
                if ((ap->dtypep() == nullptr) != (bp->dtypep() == nullptr)) return_false("dtype null");
                if (ap->dtypep() && ap->dtypep() != bp->dtypep()
                    && !ap->dtypep()->skipRefp()->similarDType(bp->dtypep()->skipRefp())) return_false("dtype");
            }
            if (!congruent(ap->op1p(), bp->op1p(), scopeAp, scopeBp, locals, rels, modes, true)) return false;
            if (!congruent(ap->op2p(), bp->op2p(), scopeAp, scopeBp, locals, rels, modes, true)) return false;
            if (!congruent(ap->op3p(), bp->op3p(), scopeAp, scopeBp, locals, rels, modes, true)) return false;
            if (!congruent(ap->op4p(), bp->op4p(), scopeAp, scopeBp, locals, rels, modes, true)) return false;
            if (!next) break;
        }
        return true;
    }
    void findSharedCode() {
        for (size_t i = 0; i < m_scopeInsts.size(); ++i) {
            m_scopeIndex[m_scopeInsts[i].scopep] = i;
            m_symIndex["(&vlSymsp->" + m_scopeInsts[i].sym + ")"] = i;
        }
        if (!V3EmitCuda::shareCode()) return;
        for (size_t c = 0; c < m_subClasses.size(); ++c) {
            std::vector<const CudaScopeInst*> insts;
            for (const CudaScopeInst& inst : m_scopeInsts) {
                if (inst.classIdx == c) insts.push_back(&inst);
            }
            if (insts.size() < 2) continue;
            // A function of a later instance is a copy if it is congruent to a function (already
            // kept) of an earlier one; V3Order numbers and splits the functions of each instance on
            // its own, so every kept function of the same kind is tried
            std::vector<std::pair<AstCFunc*, size_t>> kept;  // function, index of its scope in m_scopeInsts
            const auto kindOf = [](const AstCFunc* funcp) {
                const string::size_type pos = funcp->name().find("__TOP__");
                return pos == string::npos ? funcp->name() : funcp->name().substr(0, pos);
            };
            for (const CudaScopeInst* const instp : insts) {
                for (AstNode* nodep = m_subClasses[c].modp->stmtsp(); nodep; nodep = nodep->nextp()) {
                    AstCFunc* const funcp = VN_CAST(nodep, CFunc);
                    if (!funcp || !isDeviceFunc(funcp) || funcp->scopep() != instp->scopep || !funcp->isLoose()
                        || funcp->isStatic() || funcp->name() == "_ctor_var_reset") {
                        continue;
                    }
                    bool shared = false;
                    for (const auto& cand : kept) {
                        AstCFunc* const repp = cand.first;
                        if (repp->scopep() == funcp->scopep() || kindOf(repp) != kindOf(funcp)) continue;
                        if (repp->rtnTypeVoid() != funcp->rtnTypeVoid() || repp->slow() != funcp->slow()) continue;
                        std::map<const AstVar*, const AstVar*> locals;
                        std::set<long> rels;
                        std::map<const AstVarRef*, bool> modes;
                        if (getenv("VLCU_DEBUG_SHARE")) std::cerr << "try " << funcp->name() << " ~ " << repp->name() << "\n";
                        if (!congruent(repp, funcp, repp->scopep(), funcp->scopep(), locals, rels, modes, false)) continue;
                        for (const auto& mode : modes) m_refAbsolute[mode.first] = mode.second;
                        m_sharedAs[funcp] = repp;
                        m_sharedRep.insert(repp);
                        for (const long d : rels) {
                            std::set<size_t>& users = m_relUsed[std::make_pair(c, d)];
                            users.insert(cand.second);
                            users.insert(m_scopeIndex.at(funcp->scopep()));
                        }
                        shared = true;
                        break;
                    }
                    if (!shared) kept.emplace_back(funcp, m_scopeIndex.at(instp->scopep));
                }
            }
        }
    }
    const AstCFunc* sharedFunc(const AstCFunc* funcp) const {
        const auto it = m_sharedAs.find(funcp);
        return it == m_sharedAs.end() ? funcp : it->second;
    }
    static string relInfoName(const string& className, long distance) {
        return className + "__rel_" + (distance < 0 ? "m" + cvtToStr(-distance) : "p" + cvtToStr(distance));
    }

    // VISITORS
    using EmitCFunc::visit;
    void visit(AstConst* nodep) override {
        if (m_constInitMode) {
            EmitCConstInit::visit(nodep);  // initializer form: VlWide<N>{{...}}
        } else {
            EmitCFunc::visit(nodep);
        }
    }
    void visit(AstCFunc* nodep) override {
        // EmitCFunc::visit(AstCFunc*) (V3EmitCFunc.h:398-500) for a device function
        if (m_sharedAs.count(nodep)) return;  // another instance's copy of shared code
        VL_RESTORER(m_useSelfForThis);
        VL_RESTORER(m_cfuncp);
        VL_RESTORER(m_relSelfp);
        m_relSelfp = m_sharedRep.count(nodep) ? nodep->scopep() : nullptr;
        m_cfuncp = nodep;
        puts("\n");
        // Tiled form: the big functions (the leaves V3Order builds, V3OrderCFuncEmitter.h:137-157,
        // split by --output-split-cfuncs) are real calls, so that nvcc compiles each on its own
        // and the instances of a module share one copy; the small ones are inlined
        // Tiled form: the big functions (the leaves V3Order builds, V3OrderCFuncEmitter.h:137-157)
        // are real calls, so that ptxas compiles each on its own; the small ones are inlined.
        // With --cuda-chunk-nodes N the per-eval leaves are inlined into chunk functions of about
        // N AST nodes instead (visit(AstStmtExpr)); measured slower on the VeeR-class design
        // (more spills, worse instruction-cache behaviour), so off by default.
        bool noinline = false;
        if (m_stream && (nodep->slow() || !V3EmitCuda::chunkNodes())) {
            size_t nodes = 0;
            nodep->foreach([&](const AstNode*) { ++nodes; });
            noinline = nodes > 400;
        }
        putns(nodep, string{noinline ? "VLCU_DN" : "VLCU_D"} + " static " + nodep->rtnTypeVoid() + " "
                         + funcNameProtect(nodep) + "(" + cudaFuncArgs(nodep) + ") {\n");
        // VL_FATAL_MT leaves the function (the C++ model aborts the process there)
        puts("typedef " + nodep->rtnTypeVoid() + " VlcuFuncRet;\n");
        if (nodep->isLoose() && !nodep->isStatic()) {
            m_useSelfForThis = true;
            m_usevlSelfRef = true;
            puts("auto& vlSelfRef = *vlSelf;\n");
            if (m_stream) {
                puts("const uint32_t vlLane = vlcu_lane();  // this instance's slot in the tile\n");
            } else {
                // literal C text of the scheduler refers to vlSymsp->_vm_contextp__
                puts("const VlcuSymsShim vlSymsShim{&vlCtx};\n");
                puts("const VlcuSymsShim* const vlSymsp VL_ATTR_UNUSED = &vlSymsShim;\n");
            }
        }
        for (AstNode* subnodep = nodep->argsp(); subnodep; subnodep = subnodep->nextp()) {
            if (AstVar* const varp = VN_CAST(subnodep, Var)) {
                if (varp->isFuncReturn()) emitVarDecl(varp);
            }
        }
        // EmitCFunc prints `const uint64_t __VscopeHash = ...` once per function and remembers it in
        // a private flag that only ITS visit(AstCFunc) resets; with several classes each
        // _ctor_var_reset needs its own, so declare it here and let the body's own (if it comes)
        // shadow it in an inner block
        bool scopeHash = false;
        if (m_stream && nodep->isLoose() && !nodep->isStatic() && nodep->slow()) {
            nodep->foreach([&](const AstCReset*) { scopeHash = true; });
        }
        if (scopeHash) puts("const uint64_t __VscopeHash VL_ATTR_UNUSED = VL_MURMUR64_HASH(vlSelf->vlNamep);\n{\n");
        // Tiled form, leaves (straight-line code of one scope, no calls): gather every scalar
        // member the function touches into a local first, scatter the written ones at the end.
        // The loads are then independent of every store and of each other, so they are all in
        // flight at once -- with one warp or less per scheduler nothing else hides HBM latency.
        VL_RESTORER(m_caching);
        m_cached.clear();
        m_caching = false;
        if (m_stream && V3EmitCuda::leafCaching() && !scopeHash && leafCanCache(nodep)) {
            collectCached(nodep);
            m_caching = !m_cached.empty();
        }
        if (m_caching) {
            std::vector<const CudaCached*> order;
            for (const auto& it : m_cached) order.push_back(&it.second);
            std::sort(order.begin(), order.end(), [](const CudaCached* ap, const CudaCached* bp) {
                return std::stoul(ap->local.substr(4)) < std::stoul(bp->local.substr(4));
            });
            putsDecoration(nodep, "// Gather\n");
            m_caching = false;
            for (const CudaCached* const cp : order) {
                puts(cp->ctype + " " + cp->local);
                if (cp->load) {
                    puts(" = ");
                    iterateConst(cp->refp);
                }
                puts(";\n");
            }
            m_caching = true;
        }
        if (nodep->varsp()) {
            putsDecoration(nodep, "// Locals\n");
            iterateAndNextConstNull(nodep->varsp());
        }
        VL_RESTORER_CLEAR(m_covHitWindows);
        VL_RESTORER_CLEAR(m_covHitRepeated);
        if (m_coverBins && (m_stream || !V3EmitCuda::coverPerInstance()) && nodep->isLoose() && !nodep->isStatic()
            && !nodep->slow() && nodep->rtnTypeVoid() == "void") {
            // (a bin with several points in one function -- V3Unroll's copies of a loop body -- can
            // be passed several times in one run: those keep the direct bump)
            std::map<int, int> points;
            nodep->foreach([&](AstCoverInc* incp) {
                if (covHitEligible(incp)) ++points[coverageBinNum(incp->declp())];
            });
            for (const auto& it : points) {
                if (it.second > 1) {
                    m_covHitRepeated.insert(it.first);
                } else {
                    m_covHitWindows.insert(it.first / 32);
                }
            }
            for (const int window : m_covHitWindows) puts("uint32_t vlCovHit" + cvtToStr(window) + " = 0U;\n");
        }
        if (nodep->stmtsp()) {
            putsDecoration(nodep, "// Body\n");
            iterateAndNextConstNull(nodep->stmtsp());
        }
        for (const int window : m_covHitWindows) {
            puts("vlcu_cov_flush_hits(vlSelf->__Vcoverage + " + cvtToStr(window * 32) + ", vlCovHit" + cvtToStr(window)
                 + ");\n");
        }
        std::vector<const CudaCached*> prefetches;
        if (m_caching) {
            for (const auto& it : m_cached) {
                if (it.second.load) prefetches.push_back(&it.second);
            }
        }
        if (m_caching) {
            putsDecoration(nodep, "// Scatter\n");
            m_caching = false;
            std::vector<const CudaCached*> order;
            for (const auto& it : m_cached) {
                if (it.second.store) order.push_back(&it.second);
            }
            std::sort(order.begin(), order.end(), [](const CudaCached* ap, const CudaCached* bp) {
                return std::stoul(ap->local.substr(4)) < std::stoul(bp->local.substr(4));
            });
            for (const CudaCached* const cp : order) {
                iterateConst(cp->refp);
                puts(" = static_cast<" + cp->memType + ">(" + cp->local + ");\n");
            }
        }
        if (scopeHash) puts("}\n");
        puts("}\n");
        // ... and its companion: the same gather as prefetches, which the CALLER issues one call
        // ahead (visit(AstCCall)), so that the lines arrive while the previous leaf computes
        if (m_prefetchable.count(nodep)) {
            puts("VLCU_D static void " + funcNameProtect(nodep) + "__prefetch("
                 + EmitCUtil::prefixNameProtect(EmitCParentModule::get(nodep)) + "* vlSelf, " + m_symsName + "* vlSymsp) {\n");
            puts("auto& vlSelfRef = *vlSelf;\n");
            puts("const uint32_t vlLane = vlcu_lane();\n(void)vlSymsp; (void)vlSelfRef;\n");
            std::sort(prefetches.begin(), prefetches.end(), [](const CudaCached* ap, const CudaCached* bp) {
                return std::stoul(ap->local.substr(4)) < std::stoul(bp->local.substr(4));
            });
            for (const CudaCached* const cp : prefetches) {
                puts("vlcu_prefetch(&");
                iterateConst(cp->refp);
                puts(");\n");
            }
            puts("}\n");
        }
        m_cached.clear();
        m_usevlSelfRef = false;
    }
    void visit(AstCCall* nodep) override {
        // EmitCFunc::visit(AstCCall*) + emitCCallArgs (V3EmitCFunc.h:753-779, V3EmitCFunc.cpp:368-395)
        const AstCFunc* const funcp = sharedFunc(nodep->funcp());
        UASSERT_OBJ(funcp->isLoose() || funcp->isStatic(), nodep, "--cuda: call to a proper method");
        putnbs(nodep, "");
        // Tiled form: issue the NEXT leaf's gather as prefetches before this call
        const AstCCall* nextCallp = nullptr;
        if (m_stream && VN_IS(nodep->backp(), StmtExpr) && nodep->backp()->nextp()) {
            if (const AstStmtExpr* const nextp = VN_CAST(nodep->backp()->nextp(), StmtExpr)) {
                nextCallp = VN_CAST(nextp->exprp(), CCall);
                if (nextCallp && !m_prefetchable.count(sharedFunc(nextCallp->funcp()))) nextCallp = nullptr;
            }
        }
        if (nextCallp) {
            puts("(" + funcNameProtect(sharedFunc(nextCallp->funcp())) + "__prefetch("
                 + nextCallp->selfPointerProtect(m_useSelfForThis) + ", vlSymsp), ");
        }
        putns(nodep, funcNameProtect(funcp));
        puts("(");
        bool comma = false;
        if (funcp->isLoose() && !funcp->isStatic()) {
            if (m_stream) {
                // the callee's object: `vlSelf` or `(&vlSymsp->TOP__scope)` (V3Descope.cpp:95-125)
                puts(nodep->selfPointerProtect(m_useSelfForThis) + ", vlSymsp, vlCtx");
            } else {
                puts("vlSelf, vlCtx");
            }
            comma = true;
        }
        putCommaIterateNext(nodep->argsp(), comma);
        puts(")");
        if (nextCallp) puts(")");
    }
    // A run of consecutive leaf calls (the body of an `if (triggered)` in _eval_nba / _eval_ico /
    // _eval_act, V3OrderCFuncEmitter.h:86-113) is cut into chunks, each a real function with its
    // leaves inlined -- V3Order's --output-split-cfuncs for nvcc: ptxas sees functions of bounded
    // size, the ABI's register saves happen once per chunk instead of once per leaf, and inside a
    // chunk the next leaf's gather overlaps the previous leaf's arithmetic.
    static const AstCCall* callOfStmt(const AstNode* stmtp) {
        const AstStmtExpr* const sep = VN_CAST(stmtp, StmtExpr);
        const AstCCall* const callp = sep ? VN_CAST(sep->exprp(), CCall) : nullptr;
        if (!callp || !callp->funcp()->isLoose() || callp->funcp()->isStatic() || callp->argsp()) return nullptr;
        return callp;
    }
    static size_t funcNodes(const AstCFunc* funcp) {
        size_t nodes = 0;
        funcp->foreach([&](const AstNode*) { ++nodes; });
        return nodes;
    }
    std::set<const AstNode*> m_chunked;  // statements already printed inside a chunk
    uint32_t m_chunkCount = 0;
    void visit(AstStmtExpr* nodep) override {
        if (m_chunked.count(nodep)) return;
        const bool candidate = m_stream && V3EmitCuda::chunkNodes() && m_cfuncp && !m_cfuncp->slow() && !m_inChunk
                               && callOfStmt(nodep);
        if (candidate) {
            // the run of call statements starting here
            std::vector<AstNode*> run;
            size_t total = 0;
            for (AstNode* stmtp = nodep; stmtp && callOfStmt(stmtp); stmtp = stmtp->nextp()) {
                run.push_back(stmtp);
                total += funcNodes(callOfStmt(stmtp)->funcp());
            }
            if (run.size() >= 2 && total > V3EmitCuda::chunkNodes()) {
                size_t i = 0;
                while (i < run.size()) {
                    size_t nodes = 0, j = i;
                    while (j < run.size() && (j == i || nodes + funcNodes(callOfStmt(run[j])->funcp()) <= V3EmitCuda::chunkNodes())) {
                        nodes += funcNodes(callOfStmt(run[j])->funcp());
                        ++j;
                    }
                    const string name = "VlChunk" + cvtToStr(m_chunkCount++);
                    puts("struct " + name + " final {\nVLCU_DN static void run(" + cudaFuncArgs(m_cfuncp) + ") {\n");
                    puts("auto& vlSelfRef VL_ATTR_UNUSED = *vlSelf;\n");
                    {
                        VL_RESTORER(m_inChunk);
                        m_inChunk = true;
                        for (size_t k = i; k < j; ++k) {
                            EmitCFunc::visit(VN_AS(run[k], StmtExpr));
                            m_chunked.insert(run[k]);
                        }
                    }
                    puts("}\n};\n" + name + "::run(vlSelf, vlSymsp, vlCtx);\n");
                    i = j;
                }
                return;
            }
        }
        EmitCFunc::visit(nodep);
    }
    // Line, branch and expression points with per-warp counters: a point passed sets bit (bin % 32)
    // of a function-local mask for its window of 32 bins (one LOP3, nothing goes to memory), and
    // the masks are handed to the counters when the function ends (vlcu_cov_flush_hits: counted
    // over the warp, one add per bin).  Not inside loops (a point could be passed twice) and not
    // in functions that return a value.
    std::set<int> m_covHitWindows;  // of the function being printed; empty: points bump directly
    std::set<int> m_covHitRepeated;  // bins with several points in the function (an unrolled loop)
    int m_loopDepth = 0;
    bool covHitEligible(AstCoverInc* nodep) {
        return VN_IS(nodep->declp(), CoverOtherDecl) && coverageUsesLocalCounter(nodep->declp());
    }
    void visit(AstCoverInc* nodep) override {
        if (!m_covHitWindows.empty() && !m_loopDepth && covHitEligible(nodep)
            && !m_covHitRepeated.count(coverageBinNum(nodep->declp()))) {
            const int bin = coverageBinNum(nodep->declp());
            putns(nodep, "vlCovHit" + cvtToStr(bin / 32) + " |= " + cvtToStr(1U << (bin % 32)) + "U;\n");
            return;
        }
        EmitCFunc::visit(nodep);
    }
    void visit(AstLoop* nodep) override {
        ++m_loopDepth;
        EmitCFunc::visit(nodep);
        --m_loopDepth;
    }
    void visit(AstNodeIf* nodep) override {
        // A toggle point, IF(orig ^ change) { INC; change = orig; } (V3Clock.cpp:69-97), with per-tile
        // counters: the lanes of the warp take it together (vlcu_cov_any) so that the bump can work
        // across the warp; a lane that saw no change bumps nothing and rewrites the same shadow value.
        const AstCoverInc* const incp = VN_CAST(nodep->thensp(), CoverInc);
        if ((m_stream || !V3EmitCuda::coverPerInstance()) && incp && incp->toggleExprp() && !nodep->elsesp() && VN_IS(incp->nextp(), Assign)
            && !incp->nextp()->nextp()) {
            putns(nodep, "if (vlcu_cov_any(");
            iterateAndNextConstNull(nodep->condp());
            puts(")) {\n");
            iterateAndNextConstNull(nodep->thensp());
            puts("}\n");
            return;
        }
        EmitCFunc::visit(nodep);
    }
    bool m_inChunk = false;
    static string cacheKey(const AstVarRef* refp) { return refp->selfPointer().asString() + "." + refp->varp()->name(); }
    void visit(AstVarRef* nodep) override {
        if (m_caching) {
            const auto it = m_cached.find(cacheKey(nodep));
            if (it != m_cached.end()) {
                putns(nodep, it->second.local);
                return;
            }
        }
        bool isAbs = false;
        const long distance = m_relSelfp ? relDistance(nodep->selfPointer(), m_relSelfp, isAbs) : 0;
        const auto absIt = m_refAbsolute.find(nodep);
        if (isAbs && !nodep->varp()->isFuncLocal() && !(absIt != m_refAbsolute.end() && absIt->second)) {
            // shared code: the other scope, at its fixed distance from this function's own
            const CudaScopeInst& target = m_scopeInsts[m_scopeIndex.at(m_relSelfp) + distance];
            const string selfClass = m_subClasses[m_scopeInsts[m_scopeIndex.at(m_relSelfp)].classIdx].name;
            putns(nodep, "VLCU_REL(" + m_subClasses[target.classIdx].name + ", " + relInfoName(selfClass, distance)
                             + ")." + nodep->varp()->nameProtect());
        } else {
            EmitCFunc::visit(nodep);
        }
        // Tiled form: a member is one slot per instance of the tile
        if (m_laneVars.count(nodep->varp())) {
            puts(".v[vlLane]");
        } else if (m_laneArrayVars.count(nodep->varp())) {
            puts(".lane(vlLane)");
        }
    }
    void visit(AstDisplay* nodep) override {
        // $display/$write/$error text: the instance logs (format id, arguments); the host formats
        // it (vlcu_messages).  %m is resolved here, as displayNode passes it (V3EmitCFunc.cpp:310-317)
        string text = nodep->fmtp()->text();
        if (nodep->addNewline()) text += "\n";
        if (const AstScopeName* const scopenamep = nodep->fmtp()->scopeNamep()) {
            const string scope = "TOP." + scopenamep->scopePrettySymName();
            string::size_type pos;
            while ((pos = text.find("%m")) != string::npos) text.replace(pos, 2, scope);
        }
        size_t id = 0;
        for (; id < m_messages.size(); ++id) {
            if (m_messages[id] == text) break;
        }
        if (id == m_messages.size()) m_messages.push_back(text);
        int nargs = 0;
        for (AstNode* argp = nodep->fmtp()->exprsp(); argp; argp = argp->nextp()) ++nargs;
        if (nargs > 8) nodep->v3warn(E_UNSUPPORTED, "Unsupported: --cuda: more than 8 $display arguments");
        putns(nodep, "vlCtx.writef(" + cvtToStr(id) + "U, " + cvtToStr(nargs) + "U");
        for (AstNode* argp = nodep->fmtp()->exprsp(); argp; argp = argp->nextp()) {
            AstSFormatArg* const fargp = VN_CAST(argp, SFormatArg);
            AstNode* const subargp = fargp ? fargp->exprp() : argp;
            puts(", vlcu_msg_arg(");
            iterateConst(subargp);
            puts(")");
        }
        puts(");\n");
    }
    void visit(AstDumpCtl* nodep) override {
        putsDecoration(nodep, "// $dumpfile/$dumpvars: waveforms are written on the host (vlcu_vcd_write)\n");
    }
    void visit(AstTestPlusArgs* nodep) override {
        const AstConst* const constp = VN_CAST(nodep->searchp(), Const);
        if (!constp) {
            nodep->v3warn(E_UNSUPPORTED, "Unsupported: --cuda: $test$plusargs of a non-constant string");
            puts("0U");
            return;
        }
        const string text = constp->num().toString();
        size_t id = 0;
        for (; id < m_plusargs.size(); ++id) {
            if (m_plusargs[id] == text) break;
        }
        if (id == m_plusargs.size()) {
            // (VlcuContext::m_plusargs is one bit per string)
            if (m_plusargs.size() >= 32) nodep->v3warn(E_UNSUPPORTED, "Unsupported: --cuda: more than 32 $test$plusargs strings");
            m_plusargs.push_back(text);
        }
        putns(nodep, "static_cast<IData>(vlCtx.testPlusArg(" + cvtToStr(id) + "U))");
    }
    // _configure_coverage: each registration becomes a row of the host coverage table
    void visit(AstCoverOtherDecl* nodep) override {
        if (!m_coverTableMode) return;
        puts("t.other(" + cvtToStr(m_coverBinBase + coverageBinNum(nodep)) + ", ");
        putsQuoted(protect(nodep->fileline()->filename()));
        puts(", " + cvtToStr(nodep->fileline()->lineno()) + ", "
             + cvtToStr(nodep->offset() + nodep->fileline()->firstColumn()) + ", ");
        putsQuoted(m_coverHierPrefix + (!nodep->hier().empty() ? "." : "") + nodep->hier());
        puts(", ");
        putsQuoted(nodep->page());
        puts(", ");
        putsQuoted(nodep->comment());
        puts(", ");
        putsQuoted(nodep->linescov());
        puts(");\n");
    }
    void visit(AstCoverToggleDecl* nodep) override {
        if (!m_coverTableMode) return;
        puts("t.toggle(" + cvtToStr(m_coverBinBase + coverageBinNum(nodep)) + ", " + cvtToStr(nodep->range().right()) + ", "
             + cvtToStr(nodep->range().left()) + ", " + cvtToStr(nodep->range().ranged()) + ", ");
        putsQuoted(protect(nodep->fileline()->filename()));
        puts(", " + cvtToStr(nodep->fileline()->lineno()) + ", " + cvtToStr(nodep->fileline()->firstColumn())
             + ", ");
        putsQuoted(m_coverHierPrefix + (!nodep->hier().empty() ? "." : "") + nodep->hier());
        puts(", ");
        putsQuoted(nodep->page());
        puts(", ");
        putsQuoted(nodep->comment());
        puts(");\n");
    }

    // ---- trace declarations (trace_init_top and the functions it calls) -------------------
    // Flattened: a declaration with a dtype function (V3EmitCImp.cpp:647-652) is expanded at the
    // call site, with the name and base code the call passes ("name", "c+code")
    void traceRows(const AstNode* stmtsp, const string& dtypeName, uint32_t codeBase) {
        for (const AstNode* nodep = stmtsp; nodep; nodep = nodep->nextp()) {
            if (const AstTracePushPrefix* const pushp = VN_CAST(nodep, TracePushPrefix)) {
                // an unquoted prefix is the `name` parameter of a dtype function
                const string name = pushp->quotedPrefix() ? pushp->prefix() : dtypeName;
                m_traceRows.push_back("{VLCU_TRACE_PUSH, 0, " + quoted(name) + ", 0, VLCU_PREFIX_"
                                      + pushp->prefixType().ascii() + ", -1, " + cvtToStr(pushp->left()) + ", "
                                      + cvtToStr(pushp->right()) + "}");
            } else if (VN_IS(nodep, TracePopPrefix)) {
                m_traceRows.push_back("{VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}");
            } else if (const AstTraceDecl* const declp = VN_CAST(nodep, TraceDecl)) {
                if (const AstCCall* const callp = declp->dtypeCallp()) {
                    // the call's arguments are `tracep, "<name>"` (V3TraceDecl.cpp:487)
                    string name = declp->showname();
                    const string& args = callp->argTypes();
                    const string::size_type q0 = args.find('"');
                    const string::size_type q1 = q0 == string::npos ? q0 : args.find('"', q0 + 1);
                    if (q1 != string::npos) name = args.substr(q0 + 1, q1 - q0 - 1);
                    traceRows(callp->funcp()->stmtsp(), name, codeBase + declp->code());
                } else {
                    traceDeclRow(declp, codeBase);
                }
            } else if (const AstStmtExpr* const stmtp = VN_CAST(nodep, StmtExpr)) {
                if (const AstCCall* const callp = VN_CAST(stmtp->exprp(), CCall)) {
                    // a dtype call statement was expanded at its declaration (dtypeCallp) above
                    if (callp->argTypes().find('"') != string::npos) continue;
                    traceRows(callp->funcp()->stmtsp(), dtypeName, codeBase);
                }
            } else if (const AstCCall* const callp = VN_CAST(nodep, CCall)) {
                traceRows(callp->funcp()->stmtsp(), dtypeName, codeBase);
            }
        }
    }
    void traceDeclRow(const AstTraceDecl* declp, uint32_t codeBase) {
        const VNumRange& bits = declp->bitRange();
        const int msb = bits.ranged() ? bits.left() : 0;
        const int lsb = bits.ranged() ? bits.right() : 0;
        const uint32_t width = declp->widthMin();
        const int elements = declp->arrayRange().ranged() ? declp->arrayRange().elements() : 1;
        for (int i = 0; i < elements; ++i) {
            const uint32_t code = codeBase + declp->code() + i * declp->widthWords();
            const int elem = declp->arrayRange().ranged() ? (i + declp->arrayRange().lo()) : -1;
            m_traceRows.push_back("{VLCU_TRACE_SIG, " + cvtToStr(bits.ranged() ? 1 : 0) + ", "
                                  + quoted(declp->showname()) + ", " + cvtToStr(code) + ", " + cvtToStr(width)
                                  + ", " + cvtToStr(elem) + ", " + cvtToStr(msb) + ", " + cvtToStr(lsb) + "}");
            m_traceWords = std::max(m_traceWords, code + (width + 31) / 32);
        }
    }
    // Trace sampling.  The model's trace_full_* / trace_const_* functions (V3Trace.cpp:676-775)
    // become device functions that store every traced value at its trace code:
    //   sub function    f(vlSelf, bufp)                       -> f(vlSelf, vlCtx, valsp)
    //   dtype function  f(vlSelf, bufp, offset, __VdtypeVar)  -> template f(..., valsp, offset, var)
    static bool isDtypeTraceFunc(const AstCFunc* funcp) {
        return funcp->argTypes().find("uint32_t offset") != string::npos;
    }
    static bool isSampledTraceFunc(const AstCFunc* funcp) {
        if (!funcp->isTrace()) return false;
        return funcp->name().compare(0, 11, "trace_full_") == 0 || funcp->name().compare(0, 12, "trace_const_") == 0;
    }
    bool m_inDtypeTraceFunc = false;
    void visit(AstTraceInc* nodep) override {
        // EmitCTrace::emitTraceChangeOne / emitTraceValue (V3EmitCImp.cpp:741-810)
        const AstTraceDecl* const declp = nodep->declp();
        const VNumRange& arrayRange = declp->arrayRange();
        const int elements = arrayRange.ranged() ? arrayRange.elements() : 1;
        for (int i = 0; i < elements; ++i) {
            const uint32_t code = declp->code() + i * declp->widthWords();
            const string at = string{"valsp + "} + (m_inDtypeTraceFunc ? "offset + " : "") + cvtToStr(code);
            int arrayindex = i;
            if (arrayRange.ranged() && !arrayRange.ascending()) arrayindex = arrayRange.elements() - 1 - i;
            const string index = arrayRange.ranged() ? "[" + cvtToStr(arrayindex) + "]" : "";
            if (nodep->isWide()) {
                putns(nodep, "vlcu_trace_set_w(" + at + ", " + cvtToStr(declp->widthWords()) + ", (");
            } else {
                putns(nodep, "vlcu_trace_set(" + at + ", static_cast<" + (nodep->isQuad() ? "QData" : "IData") + ">(");
            }
            iterateConst(nodep->valuep());
            puts(index + "));\n");
        }
    }
    // A call from one trace function to another: "bufp[, ...]" + expression arguments
    void emitTraceCall(const AstCCall* callp) {
        string rest = callp->argTypes();
        if (rest.compare(0, 4, "bufp") == 0) rest = rest.substr(4);
        puts(funcNameProtect(callp->funcp()) + "(vlSelf, vlCtx, valsp" + rest);
        for (AstNode* argp = callp->argsp(); argp; argp = argp->nextp()) {
            puts(", ");
            iterateConst(argp);
        }
        puts(");\n");
    }
    void emitTraceFunc(AstCFunc* funcp) {
        VL_RESTORER(m_cfuncp);
        VL_RESTORER(m_useSelfForThis);
        VL_RESTORER(m_inDtypeTraceFunc);
        m_cfuncp = funcp;
        m_useSelfForThis = true;
        m_usevlSelfRef = true;
        m_inDtypeTraceFunc = isDtypeTraceFunc(funcp);
        puts("\n");
        if (m_inDtypeTraceFunc) puts("template <class T_DtypeVar>\n");
        puts("VLCU_D static void " + funcNameProtect(funcp) + "(" + m_rootName
             + "* vlSelf, VlcuContext& vlCtx, uint32_t* valsp"
             + (m_inDtypeTraceFunc ? ", uint32_t offset, const T_DtypeVar& __VdtypeVar" : "") + ") {\n");
        puts("auto& vlSelfRef VL_ATTR_UNUSED = *vlSelf;\n");
        if (funcp->varsp()) iterateAndNextConstNull(funcp->varsp());  // wide temporaries
        for (AstNode* stmtp = funcp->stmtsp(); stmtp; stmtp = stmtp->nextp()) {
            const AstStmtExpr* const sep = VN_CAST(stmtp, StmtExpr);
            const AstCCall* const callp = sep ? VN_CAST(sep->exprp(), CCall) : nullptr;
            if (callp && callp->funcp()->isTrace()) {
                emitTraceCall(callp);
            } else if (VN_IS(stmtp, CStmt)) {
                // literal host text (the `oldp` pointer of the trace buffer): not needed
            } else {
                // AstTraceInc (store by code), or a statement computing a wide temporary for one
                iterateConst(stmtp);
            }
        }
        m_usevlSelfRef = false;
        puts("}\n");
    }
    // Dump order of the trace codes: the statements in execution order, through the calls
    void traceOrderWalk(const AstCFunc* funcp, uint32_t base) {
        for (const AstNode* stmtp = funcp->stmtsp(); stmtp; stmtp = stmtp->nextp()) {
            if (const AstTraceInc* const incp = VN_CAST(stmtp, TraceInc)) {
                const AstTraceDecl* const declp = incp->declp();
                const int elements = declp->arrayRange().ranged() ? declp->arrayRange().elements() : 1;
                for (int i = 0; i < elements; ++i) m_traceOrder.push_back(base + declp->code() + i * declp->widthWords());
            } else if (const AstStmtExpr* const sep = VN_CAST(stmtp, StmtExpr)) {
                const AstCCall* const callp = VN_CAST(sep->exprp(), CCall);
                if (!callp || !callp->funcp()->isTrace()) continue;
                uint32_t calleeBase = base;  // "bufp" / "bufp, offset, __VdtypeVar": same base
                const string& args = callp->argTypes();
                const string::size_type comma = args.find(", ");
                if (comma != string::npos && std::isdigit(args[comma + 2])) {
                    calleeBase = std::stoul(args.substr(comma + 2));  // "bufp, <code>"
                }
                traceOrderWalk(callp->funcp(), calleeBase);
            }
        }
    }
    void emitTraceSample() {
        // Sub and dtype functions first (callees), in AST order
        std::vector<AstCFunc*> tops;
        for (AstNode* nodep = m_topModp->stmtsp(); nodep; nodep = nodep->nextp()) {
            AstCFunc* const funcp = VN_CAST(nodep, CFunc);
            if (!funcp || !isSampledTraceFunc(funcp)) continue;
            if (funcp->isStatic()) {  // top functions take `void* voidSelf` and only call sub functions
                tops.push_back(funcp);
            } else {
                emitTraceFunc(funcp);
            }
        }
        puts("\n// tfp->dump(): every traced value stored at its trace code.  Full functions, then the\n");
        puts("// constants, as VerilatedTrace::dump runs them the first time (verilated_trace_imp.h:348-358)\n");
        puts("VLCU_D static void trace_sample(Root* vlSelf, VlcuContext& vlCtx, uint32_t* valsp) {\n");
        puts("(void)vlSelf; (void)vlCtx; (void)valsp;\n");
        for (const char* const prefix : {"trace_full_", "trace_const_"}) {
            for (AstCFunc* const funcp : tops) {
                if (funcp->name().compare(0, strlen(prefix), prefix) != 0) continue;
                for (AstNode* stmtp = funcp->stmtsp(); stmtp; stmtp = stmtp->nextp()) {
                    if (const AstStmtExpr* const sep = VN_CAST(stmtp, StmtExpr)) {
                        if (const AstCCall* const callp = VN_CAST(sep->exprp(), CCall)) {
                            puts(funcNameProtect(callp->funcp()) + "(vlSelf, vlCtx, valsp);\n");
                        }
                    }
                }
                traceOrderWalk(funcp, 0);
            }
        }
        puts("}\n");
    }
    uint32_t m_traceWords = 0;
    std::vector<uint32_t> m_traceOrder;

    // ---- the file ---------------------------------------------------------------------------
    // How an entry point hands the root object to the model's functions
    string entryArgs() const { return m_stream ? "&vlSelf->symsp->TOP, vlSelf->symsp, vlCtx" : "vlSelf, vlCtx"; }
    void emitWrapper(const string& wrapper, const string& funcName) {
        const AstCFunc* const funcp = findFunc(funcName);
        puts("VLCU_D static void " + wrapper + "(Root* vlSelf, VlcuContext& vlCtx) {");
        if (funcp && isDeviceFunc(funcp)) {
            puts(" " + funcNameProtect(funcp) + "(" + entryArgs() + "); }\n");
        } else {
            puts(" (void)vlSelf; (void)vlCtx; }\n");
        }
    }
    void emitDeviceFuncs(const AstNodeModule* modp) {
        VL_RESTORER(m_modp);
        m_modp = modp;
        for (AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
            if (AstCFunc* const funcp = VN_CAST(nodep, CFunc)) {
                if (isDeviceFunc(funcp)) iterateConst(funcp);
            }
        }
    }
    void emitAll() {
        const string fileName = v3Global.opt.makeDir() + "/" + m_designName + ".cu";
        AstCFile* const cfilep = new AstCFile{v3Global.rootp()->fileline(), fileName};
        cfilep->support(true);
        openOutputFile(cfilep, "Batched CUDA model (one instance per thread) for include/verilated_cuda.h");
        puts("// Design '" + v3Global.rootp()->topModulep()->prettyName() + "' printed by V3EmitCuda; compile with\n");
        puts("//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -I$VERILATOR_ROOT/include\n");
        if (m_stream) {
            puts("\n// Tiled form: the model image stays in HBM (" + cvtToStr(m_scopeInsts.size() + 1)
                 + " scopes of " + cvtToStr(m_subClasses.size() + 1) + " module classes); its kernels hold\n");
            puts("// members in registers only inside a leaf (gather ... scatter) and are bound by memory latency.\n");
            puts("// 4 CTAs of 64 threads per SM = 255 registers per thread: the leaves' gathers then keep\n");
            puts("// 100+ loads in flight per warp without spilling (measured: profiles/README.md)\n");
            puts("#ifndef VLCU_MIN_BLOCKS\n#define VLCU_MIN_BLOCKS 4\n#endif\n");
            if (m_coverBins) {
                puts("// " + cvtToStr(m_coverBins) + " counters: the bump functions are real functions (-DVLCU_COV_INLINE: inlined,\n");
                puts("// a much longer compilation), and lanes that meet in them are grouped by the counter they hold\n");
                puts("#ifndef VLCU_COV_INLINE\n#define VLCU_COV_NOINLINE 1\n#endif\n#define VLCU_COV_GROUP_LANES 1\n");
            }
        } else {
            // Occupancy target of the kernels (verilated_cuda_launch.h: __launch_bounds__(64,
            // VLCU_MIN_BLOCKS), default 7 = 146 registers per thread).  A design that keeps almost
            // nothing in registers -- its state is in memory-class arrays -- is bound by memory
            // latency: compile it for 32 resident CTAs per SM (32 registers per thread, all 64
            // warp slots) so that more loads are in flight.
            uint32_t regBytes = 0;
            for (const CudaMember& m : m_members) {
                if (m.storage == CudaStorage::REG) regBytes += std::max<uint32_t>(m.bytes(), 4);
            }
            if (regBytes <= 64) {
                puts("\n// " + cvtToStr(regBytes) + " bytes of register-resident state: latency-bound, compile for full occupancy\n");
                puts("#ifndef VLCU_MIN_BLOCKS\n#define VLCU_MIN_BLOCKS 32\n#endif\n");
            }
        }
        puts("\n#include \"verilated_cuda_launch.h\"\n");
        puts("\n#include <vector>\n");
        puts("\n// Name the design registers under with vlcu_create()\n");
        puts("#ifndef VLCU_DESIGN_NAME\n#define VLCU_DESIGN_NAME " + quoted(v3Global.opt.prefix()) + "\n#endif\n");
        emitConstPool();
        if (m_stream) {
            emitTileStructs();
        } else {
            emitRootStruct();
        }

        puts("\nstruct " + m_designName + " final {\n");
        if (m_stream) {
            puts("typedef " + m_symsName + " Syms;\n");
            puts("// What the kernels hold per thread: the instance's tile and its slot in it\n");
            puts("struct Root final {\nSyms* symsp;\nuint32_t lane;\n};\n");
            puts("static constexpr bool kTiled = true;\n");
            puts("enum : uint32_t {\n");
            puts("kStateBytes = sizeof(Syms) / VLCU_TILE,  // S_persist\n");
        } else {
            puts("typedef " + m_rootName + " Root;\n");
            puts("static constexpr bool kTiled = false;\n");
            // Slot table
            puts("// Structure-of-arrays slot table: byte offset per instance of every persistent member\n");
            puts("enum : uint32_t {\n");
            for (const CudaMember& m : m_members) puts("OFF_" + m.name + " = " + cvtToStr(m.offset) + ",\n");
            puts("kStateBytes = " + cvtToStr(m_stateBytes) + ",  // S_persist\n");
        }
        puts("kTouchedReadBytes = " + cvtToStr(m_touchedRead) + ",  // S_state_rd: persistent bytes one eval can read\n");
        puts("kTouchedWriteBytes = " + cvtToStr(m_touchedWrite) + ",  // S_state_wr\n");
        puts("};\n");
        // Storage of the arrays that are not in registers
        puts("struct Locals final {\n");
        for (const CudaMember& m : m_members) {
            if (m.storage == CudaStorage::LOCAL) puts(m.elemDecl + " " + m.name + "[" + cvtToStr(m.depth) + "];\n");
        }
        for (const auto& q : m_nbaQueues) puts(q.second + "::Store " + q.first + ";\n");
        puts("};\n");
        puts("struct Shared final {\n");
        for (const CudaMember& m : m_members) {
            if (m.storage == CudaStorage::SHARED) {
                puts(m.elemDecl + " " + m.name + "[" + cvtToStr(m.depth) + "][VLCU_BLOCK];\n");
            }
        }
        puts("};\n");
        puts("VLCU_D static void bind(Root& s, Locals& l, Shared& sh, const VlcuState& st, uint64_t i) {\n");
        puts("(void)l; (void)sh; (void)st; (void)i;\n");
        if (m_stream) {
            puts("s.symsp = reinterpret_cast<Syms*>(st.base) + i / VLCU_TILE;\n");
            puts("s.lane = static_cast<uint32_t>(i % VLCU_TILE);\n");
            puts("}\n");
            puts("// (nothing to load or store: the functions work on the image in place)\n");
            puts("VLCU_D static void load(Root&, const VlcuState&, uint64_t) {}\n");
            puts("VLCU_D static void store(const Root&, const VlcuState&, uint64_t) {}\n");
        } else {
            puts("s.vlNamep = \"TOP\";\n");
            for (const auto& q : m_nbaQueues) puts("s." + q.first + ".bind(&l." + q.first + ");\n");
            for (const CudaMember& m : m_members) {
                if (m.storage == CudaStorage::LOCAL) puts("s." + m.name + ".bind(l." + m.name + ");\n");
                if (m.storage == CudaStorage::SHARED) puts("s." + m.name + ".bind(sh." + m.name + ", threadIdx.x);\n");
                if (m.storage == CudaStorage::GLOBAL || m.storage == CudaStorage::COVER
                    || m.storage == CudaStorage::INTERLEAVED) {
                    puts("s." + m.name + ".bind(st, OFF_" + m.name + ", i);\n");
                }
            }
            puts("}\n");
            emitLoadStore(true);
            emitLoadStore(false);
        }
        emitPortRows();
        if (!m_stream) {
            if (const AstCFunc* const funcp = findFunc("trace_init_top")) traceRows(funcp->stmtsp(), "", 0);
        }
        puts("enum : uint32_t { " + m_sizes + ", kTraceWords = " + cvtToStr(m_traceWords) + " };\n");

        // Device functions
        if (m_stream && V3EmitCuda::leafCaching() && V3EmitCuda::leafPrefetch()) {
            // which leaves gather (and so have a prefetch companion): known before any call is printed
            const auto scan = [this](const AstNodeModule* modp) {
                for (AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
                    AstCFunc* const funcp = VN_CAST(nodep, CFunc);
                    if (!funcp || !isDeviceFunc(funcp) || !leafCanCache(funcp)) continue;
                    m_cached.clear();
                    collectCached(funcp);
                    bool loads = false;
                    for (const auto& it : m_cached) loads |= it.second.load;
                    if (loads) m_prefetchable.insert(funcp);
                    m_cached.clear();
                }
            };
            for (const CudaClass& cls : m_subClasses) scan(cls.modp);
            scan(m_topModp);
        }
        for (const CudaClass& cls : m_subClasses) {
            if (m_stream) emitDeviceFuncs(cls.modp);
        }
        emitDeviceFuncs(m_topModp);
        if (m_stream) {
            // (waveforms of a tiled design: not yet)
            puts("\nVLCU_D static void trace_sample(Root*, VlcuContext&, uint32_t*) {}\n");
        } else {
            emitTraceSample();
        }
        // Entry points of the launch templates = what Vtop::eval_step() calls
        // (EmitCModel::emitStandardMethods1, V3EmitCModel.cpp:403-437)
        puts("\n");
        // Construction: the member initialiser `uint32_t __Vcoverage[n]{}` of the C++ class
        // (V3EmitCHeaders.cpp:153-158), then Vtop___024root___ctor_var_reset
        puts("VLCU_D static void ctor_var_reset(Root* vlSelf, VlcuContext& vlCtx) {\n");
        if (m_stream) {
            // what the Syms constructor does (V3EmitCSyms.cpp:1148-1173): every scope is named
            // -- the reset values hash the name, V3EmitCFunc.cpp:668-678 -- and constructed
            puts("Syms* const vlSymsp = vlSelf->symsp;\n");
            puts("vlSymsp->TOP.vlNamep = \"TOP\";\n");
            if (m_classCoverBins[m_topModp]) puts("vlSymsp->TOP.__Vcoverage.zero(vlSelf->lane);\n");
            for (const CudaScopeInst& inst : m_scopeInsts) {
                if (m_classCoverBins[m_subClasses[inst.classIdx].modp]) {
                    puts("vlSymsp->" + inst.sym + ".__Vcoverage.zero(vlSelf->lane);\n");
                }
            }
            if (const AstCFunc* const funcp = findFunc("_ctor_var_reset")) {
                puts(funcNameProtect(funcp) + "(&vlSymsp->TOP, vlSymsp, vlCtx);\n");
            }
            for (const CudaScopeInst& inst : m_scopeInsts) {
                puts("vlSymsp->" + inst.sym + ".vlNamep = " + quoted("TOP." + inst.pretty) + ";\n");
                for (const AstNode* nodep = m_subClasses[inst.classIdx].modp->stmtsp(); nodep; nodep = nodep->nextp()) {
                    const AstCFunc* const funcp = VN_CAST(nodep, CFunc);
                    if (funcp && funcp->name() == "_ctor_var_reset") {
                        puts(funcNameProtect(sharedFunc(funcp)) + "(&vlSymsp->" + inst.sym + ", vlSymsp, vlCtx);\n");
                    }
                }
            }
        } else {
            if (m_coverBins) {
                if (V3EmitCuda::coverPerInstance()) {
                    puts("for (uint32_t b = 0; b < " + cvtToStr(m_coverBins) + "U; ++b) vlSelf->__Vcoverage[b] = 0U;\n");
                } else {
                    puts("vlSelf->__Vcoverage.zero(" + cvtToStr(m_coverBins) + "U);\n");
                }
            }
            if (const AstCFunc* const funcp = findFunc("_ctor_var_reset")) {
                puts(funcNameProtect(funcp) + "(vlSelf, vlCtx);\n");
            }
        }
        puts("}\n");
        emitWrapper("eval_static", "_eval_static");
        emitWrapper("eval_initial", "_eval_initial");
        emitWrapper("eval_settle", "_eval_settle");
        if (m_nbaQueues.empty()) {
            emitWrapper("eval", "_eval");
        } else {
            // a full commit queue cannot grow on the device: the instance takes a fatal instead
            puts("VLCU_D static void eval(Root* vlSelf, VlcuContext& vlCtx) {\n");
            if (const AstCFunc* const funcp = findFunc("_eval")) {
                puts(funcNameProtect(funcp) + "(" + entryArgs() + ");\n");
            }
            for (const auto& q : m_nbaQueues) {
                puts("if (VL_UNLIKELY(vlSelf->" + q.first + ".overflowed())) vlCtx.fatal(0, VLCU_FATAL_NBAQ_OVERFLOW);\n");
            }
            puts("}\n");
        }
        emitWrapper("eval_final", "_eval_final");
        puts("\nstatic const VlcuDesignInfo& info();\n");
        puts("};\n");

        emitHostTables();
        emitPortsInc();
        emitMk();
        puts("\n// The device testbenches of this design (the user's C++ main loop, on the GPU)\n");
        puts("#ifdef VLCU_TESTBENCH_INCLUDE\n#include VLCU_TESTBENCH_INCLUDE\n#endif\n");
        puts("#ifndef VLCU_TESTBENCHES  // the testbench header lists its classes: #define VLCU_TESTBENCHES , TbA, TbB\n");
        puts("#define VLCU_TESTBENCHES\n#endif\n");
        puts("VLCU_REGISTER_DESIGN_X(" + m_designName + " VLCU_TESTBENCHES);\n");
        closeOutputFile();
    }
    // <prefix>__ports.inc: the primary ports in source order as X-macro calls, for generic host
    // harnesses that drive the C++ model with the same stimulus / observation rows
    void emitPortsInc() {
        const string fileName = v3Global.opt.makeDir() + "/" + v3Global.opt.prefix() + "__ports.inc";
        std::ofstream os{fileName};
        os << "// Primary ports of " << v3Global.opt.prefix() << " in source order (printed by V3EmitCuda)\n";
        os << "// VLCU_PORT_IN / VLCU_PORT_OUT / VLCU_PORT_INOUT (member name, width in bits, 32-bit row words)\n";
        for (const CudaMember* const mp : portsInSourceOrder()) {
            os << (mp->kind == 1 ? "VLCU_PORT_OUT(" : mp->kind == 2 ? "VLCU_PORT_INOUT(" : "VLCU_PORT_IN(")
               << mp->name << ", " << mp->width << ", " << rowWords(*mp) << ")\n";
        }
    }
    // <prefix>__cuda.mk: what V3EmitMk's <prefix>.mk is for the C++ model -- the design's
    // settings, then the generic nvcc rules of include/verilated_cuda.mk
    void emitMk() {
        const string prefix = v3Global.opt.prefix();
        V3OutMkFile of{v3Global.opt.makeDir() + "/" + prefix + "__cuda.mk"};
        of.putsHeader();
        of.puts("# DESCRIPTION: Verilator output: Makefile for building the batched CUDA model\n");
        of.puts("#\n");
        of.puts("# Execute this makefile from the object directory:\n");
        of.puts("#    make -f " + prefix + "__cuda.mk [VLCU_TESTBENCH=my_tb.h] [VLCU_DESIGN=name]\n");
        of.puts("# It builds lib" + prefix + "__cuda.so: this design's device code + the batch runtime behind\n");
        of.puts("# the C ABI of verilated_cuda_abi.h.\n");
        of.puts("\n### Constants...\n");
        of.puts("# Path to Verilator kit (from $VERILATOR_ROOT)\n");
        of.puts("VERILATOR_ROOT ?= " + V3Options::getenvVERILATOR_ROOT() + "\n");
        of.puts("# Where verilated_cuda*.h, verilated_cuda.cu and verilated_cuda.mk live\n");
        of.puts("VLCU_ROOT ?= $(VERILATOR_ROOT)\n");
        of.puts("\n### Switches...\n");
        of.puts("# Generated module classes prefix (from --prefix)\n");
        of.puts("VM_PREFIX = " + prefix + "\n");
        of.puts("# Name the design registers under (vlcu_create(design, ...))\n");
        of.puts("VLCU_DESIGN ?= " + prefix + "\n");
        of.puts("# Header with the design's device testbenches (optional; see verilated_cuda_launch.h)\n");
        of.puts("VLCU_TESTBENCH ?=\n");
        of.puts("# Coverage counters per instance (from --coverage)\n");
        of.puts("VLCU_COVER_BINS = " + cvtToStr(m_coverBins) + "\n");
        of.puts("# Persistent bytes per instance\n");
        of.puts("VLCU_STATE_BYTES = " + cvtToStr(m_stateBytes) + "\n");
        of.puts("\n### Include the generic rules...\n");
        of.puts("include $(VLCU_ROOT)/include/verilated_cuda.mk\n");
    }
    void emitHostTables() {
        // The user's top module: first component of the internal signals' hierarchical names
        string topName = v3Global.opt.prefix();
        for (const CudaMember& m : m_members) {
            const string::size_type dot = m.hier.find('.');
            if (m.varp && dot != string::npos && m.hier[0] != '_') {
                topName = m.hier.substr(0, dot);
                break;
            }
        }
        puts("\n//======================================================================\n");
        puts("// Host tables: what Vtop.h / Vtop__Syms / __Vconfigure give the host runtime\n");
        puts("\nconst VlcuDesignInfo& " + m_designName + "::info() {\n");
        puts("typedef " + m_designName + " T;\n");
        puts("// name, hierName, kind, elemBytes, flags, elems, width, msb, lsb, soaOffset\n");
        puts("static const VlcuMemberDesc members[] = {\n");
        for (const CudaFlatMember& fm : m_flat) {
            const CudaMember& m = fm.m;
            static const char* const kinds[]
                = {"VLCU_MEMBER_IN", "VLCU_MEMBER_OUT", "VLCU_MEMBER_INOUT", "VLCU_MEMBER_STATE",
                   "VLCU_MEMBER_COVERAGE", "VLCU_MEMBER_MEMORY"};
            puts("{" + quoted(fm.tableName) + ", " + (m.varp ? quoted(m.hier) : string{"nullptr"}) + ", " + kinds[m.kind]
                 + ", " + cvtToStr(m.elemBytes) + ", " + (m.laneMajor ? "VLCU_MEMBERF_LANE_MAJOR" : "0") + ", "
                 + cvtToStr(m.elems()) + ", " + cvtToStr(m.width) + ", "
                 + cvtToStr(m.msb) + ", " + cvtToStr(m.lsb) + ", " + fm.offsetExpr + "},\n");
        }
        puts("};\n");
        // Coverage
        if (m_coverBins) {
            puts("static VlcuCoverTable t;\n");
            puts("if (t.points.empty()) {\n");
            // what __Vconfigure registers (V3EmitCImp.cpp:163-262), scope by scope
            const auto listPoints = [this](const AstNodeModule* modp, int binBase, const string& hierPrefix) {
                for (AstNode* nodep = modp->stmtsp(); nodep; nodep = nodep->nextp()) {
                    AstCFunc* const funcp = VN_CAST(nodep, CFunc);
                    if (!funcp || funcp->name() != "_configure_coverage") continue;
                    VL_RESTORER(m_coverTableMode);
                    VL_RESTORER(m_cfuncp);
                    VL_RESTORER(m_modp);
                    VL_RESTORER(m_coverBinBase);
                    VL_RESTORER_COPY(m_coverHierPrefix);
                    m_coverTableMode = true;
                    m_cfuncp = funcp;  // coverageBinNum() picks the object-local bin inside a member function
                    m_modp = modp;
                    m_coverBinBase = binBase;
                    m_coverHierPrefix = hierPrefix;
                    for (AstNode* stmtp = funcp->stmtsp(); stmtp; stmtp = stmtp->nextp()) {
                        if (VN_IS(stmtp, NodeCoverDecl)) iterateConst(stmtp);
                    }
                }
            };
            listPoints(m_topModp, 0, "");
            if (m_stream) {
                int base = m_classCoverBins[m_topModp];
                for (const CudaScopeInst& inst : m_scopeInsts) {
                    const AstNodeModule* const modp = m_subClasses[inst.classIdx].modp;
                    listPoints(modp, base, "." + inst.pretty);  // vlNamep = "TOP.<scope>"
                    base += m_classCoverBins[modp];
                }
            }
            puts("t.finish();\n");
            puts("}\n");
            if (!m_stream && !V3EmitCuda::coverPerInstance()) {
                puts("// one uint64 counter per bin for every warp of 32 instances, [warp][bin] in the block of __Vcoverage\n");
                puts("static const VlcuCoverScope coverScopes[] = {{0, " + cvtToStr(m_coverBins) + "}};\n");
            }
            if (m_stream) {
                puts("// where each scope's counters are in a tile: byte offset of counter 0 (uint64 each), counters\n");
                puts("static const VlcuCoverScope coverScopes[] = {\n");
                if (const int bins = m_classCoverBins[m_topModp]) {
                    puts("{VLCU_TILE_OFF(" + m_symsName + ", TOP, __Vcoverage), " + cvtToStr(bins) + "},\n");
                }
                for (const CudaScopeInst& inst : m_scopeInsts) {
                    if (const int bins = m_classCoverBins[m_subClasses[inst.classIdx].modp]) {
                        puts("{VLCU_TILE_OFF(" + m_symsName + ", " + inst.sym + ", __Vcoverage), " + cvtToStr(bins) + "},\n");
                    }
                }
                puts("};\n");
            }
        }
        // Messages and plusargs
        puts("static const char* const messages[] = {");
        for (const string& s : m_messages) puts(quoted(s) + ", ");
        puts("nullptr};\n");
        puts("static const char* const plusargs[] = {");
        for (const string& s : m_plusargs) puts(quoted(s) + ", ");
        puts("nullptr};\n");
        // Trace declarations
        // (rows were collected before the device code was printed, for kTraceWords)
        puts("static const VlcuTraceDecl traceDecls[] = {\n");
        for (const string& row : m_traceRows) puts(row + ",\n");
        puts("{VLCU_TRACE_POP, 0, nullptr, 0, 0, -1, 0, 0}};  // (terminator, not counted)\n");
        puts("static const uint32_t traceOrder[] = {");
        for (const uint32_t code : m_traceOrder) puts(cvtToStr(code) + ", ");
        puts("0};  // codes in dump order (+ terminator)\n");
        // verilated_save image: the members EmitCImp::emitSavableImp (V3EmitCImp.cpp:277-350)
        // would serialize, in its order, and its structure checksum
        {
            VHashSha256 hash;
            string unsupported;
            puts("static const uint32_t saveOrder[] = {");
            for (const AstNode* nodep = m_topModp->stmtsp(); nodep; nodep = nodep->nextp()) {
                const AstVar* const varp = VN_CAST(nodep, Var);
                if (!varp) continue;
                hash.insert(varp->name());
                hash.insert(varp->dtypep()->width());
                if (varp->isParam()) continue;
                if (varp->isStatic() && varp->isConst()) continue;
                if (VN_IS(varp->dtypep(), NBACommitQueueDType)) continue;
                const AstBasicDType* const basicp = varp->dtypeSkipRefp()->basicp();
                if (basicp && basicp->keyword().isMTaskState()) continue;
                const auto it = m_memberOf.find(varp);
                if (it == m_memberOf.end()) {
                    if (unsupported.empty()) unsupported = varp->name();
                    continue;
                }
                puts(cvtToStr(it->second) + ", ");
                ++m_nSaveOrder;
            }
            puts("0};  // (+ terminator)\n");
            m_saveCheckval = hash.digestUInt64();
            if (m_stream && unsupported.empty()) unsupported = "(tiled model image)";
            m_saveUnsupported = unsupported;
        }
        puts("static const VlcuDesignInfo s = {\n");
        puts("VLCU_DESIGN_NAME, " + quoted(topName) + ", members, sizeof(members) / sizeof(members[0]),\n");
        puts("T::kStateBytes, T::kInWords, T::kOutWords, T::kInBytes, T::kOutBytes, T::kTouchedReadBytes, "
             "T::kTouchedWriteBytes,\n");
        if (m_coverBins && m_stream) {
            puts("t.points.data(), static_cast<uint32_t>(t.points.size()), " + cvtToStr(m_coverBins) + ", 0,\n");
        } else if (m_coverBins) {
            puts("t.points.data(), static_cast<uint32_t>(t.points.size()), " + cvtToStr(m_coverBins)
                 + ", T::OFF___Vcoverage,\n");
        } else {
            puts("nullptr, 0, 0, 0,\n");
        }
        puts("messages, " + cvtToStr(m_messages.size()) + ", plusargs, " + cvtToStr(m_plusargs.size()) + ",\n");
        puts(m_traceRows.empty() ? "nullptr, 0, 0, nullptr, 0,\n"
                                 : "traceDecls, " + cvtToStr(m_traceRows.size()) + ", T::kTraceWords, traceOrder, "
                                       + cvtToStr(m_traceOrder.size()) + ",\n");
        {
            std::ostringstream os;
            os << "saveOrder, " << m_nSaveOrder << ", " << (v3Global.opt.trace() ? "VLCU_SAVE_ACTIVITY" : "0")
               << ", 0x" << std::hex << m_saveCheckval << "ULL, "
               << (m_saveUnsupported.empty() ? string{"nullptr"} : quoted(m_saveUnsupported)) << ", " << std::dec
               << (v3Global.rootp()->timeunit().isNone() ? -1 : std::abs(v3Global.rootp()->timeunit().powerOfTen()))
               << ", "
               << (v3Global.rootp()->timeprecision().isNone()
                       ? -1
                       : std::abs(v3Global.rootp()->timeprecision().powerOfTen()))
               << ", " << (m_stream ? "VLCU_TILE" : "0")
               << (m_coverBins && (m_stream || !V3EmitCuda::coverPerInstance())
                       ? ", coverScopes, sizeof(coverScopes) / sizeof(coverScopes[0])"
                       : ", nullptr, 0")
               << "};\n";
            puts(os.str());
        }
        puts("return s;\n}\n");
    }

public:
    EmitCudaModel()
        : m_topModp{v3Global.rootp()->topModulep()}
        , m_rootName{EmitCUtil::prefixNameProtect(v3Global.rootp()->topModulep())}
        , m_designName{v3Global.opt.prefix() + "__cuda"}
        , m_symsName{v3Global.opt.prefix() + "__Syms"} {
        m_modp = m_topModp;
        computeLayout();
        emitAll();
        V3Stats::addStat("CUDA model, persistent bytes per instance", m_stateBytes);
    }
};

}  // namespace

bool V3EmitCuda::s_enabled = false;
bool V3EmitCuda::s_layoutSelection = true;
int V3EmitCuda::s_streamMode = -1;
bool V3EmitCuda::s_leafCaching = true;
bool V3EmitCuda::s_leafPrefetch = false;
bool V3EmitCuda::s_coverPerInstance = false;
int V3EmitCuda::s_chunkNodes = 0;
bool V3EmitCuda::s_shareCode = true;
bool V3EmitCuda::s_interleaveAll = false;
int V3EmitCuda::s_gatherWords = 0;

void V3EmitCuda::emitCuda() {
    UINFO(2, __FUNCTION__ << ":");
    const EmitCParentModule emitCParentModule;  // user4p -> parent module, as V3EmitC::emitcImp does
    { const EmitCudaModel emitter; }
}
