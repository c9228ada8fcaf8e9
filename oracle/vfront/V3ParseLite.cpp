// -*- mode: C++; c-file-style: "cc-mode" -*-
//*************************************************************************
// TEST INFRASTRUCTURE (oracle side, never part of the product).
//
// V3ParseLite -- a small hand-written Verilog front end that stands in for
// the reference's bison/flex generated parser, which cannot be built in this
// image (no bison, no flex).  Everything AFTER parsing is the
// reference's own code, compiled from /root/reference/src where it lies:
// V3LinkCells, V3LinkParse, V3Width, ... V3Sched, V3Order, ... V3EmitC*.
// With this file providing the few symbols the parser normally provides
// (V3Parse::*, V3PreProc::createPreProc) the rest of verilator_bin links
// unchanged, and `oracle/_ref/verilator_lite` turns the designs of this repo
// into GENUINE Verilated C++ (the oracle) and, through src/V3EmitCuda.cpp,
// into the CUDA model.
//
// What it accepts: the synthesizable Verilog-2005/SystemVerilog subset the
// designs under verilator_b200/csrc/designs and examples/make_tracing_c use
// (ANSI module headers, nets/variables/memories/parameters, continuous and
// procedural assignments, always/always_ff/always_comb/initial, if/case/for,
// module instances with named or ordered pins, the common operators and
// system tasks, immediate assert and `cover property`).  Anything else is a
// syntax error -- this is not a general parser.
//
// What it builds: the same raw AST verilog.y's actions build for those
// constructs (node types, child order, flags and FileLine of the token the
// grammar uses); each construct cites its rule in src/verilog.y.
//*************************************************************************

#include "V3PchAstNoMT.h"

#include "V3Ast.h"
#include "V3File.h"
#include "V3Global.h"
#include "V3Number.h"
#include "V3Os.h"
#include "V3Parse.h"
#include "V3PreProc.h"
#include "V3PreShell.h"
#include "V3String.h"

#include <cctype>
#include <deque>
#include <map>
#include <fstream>
#include <set>
#include <sstream>

VL_DEFINE_DEBUG_FUNCTIONS;

//======================================================================
// Preprocessor stand-in: hands the file through unchanged, framed by the `line
// directive the real preprocessor emits (V3PreProc.cpp), so the parser knows the file name

namespace {

class PreProcLite final : public V3PreProc {
    std::string m_text;
    bool m_eof = true;

public:
    void openFile(FileLine* fl, VInFilter*, const string& filename) override {
        std::ifstream is{filename};
        if (is.fail()) {
            fl->v3error("File not found: " + filename);
            return;
        }
        std::stringstream ss;
        ss << is.rdbuf();
        V3File::addSrcDepend(filename);
        m_text = "`line 1 \"" + filename + "\" 1\n" + ss.str();
        m_eof = false;
    }
    string getline() override {
        m_eof = true;
        return std::move(m_text);
    }
    bool isEof() const override { return m_eof; }
    void insertUnreadback(const string&) override {}
    void comment(const string&) override {}
    void include(const string&) override {}
    void undef(const string&) override {}
    void define(FileLine*, const string&, const string&, const string&, bool) override {}
    string removeDefines(const string& text) override { return text; }
    void dumpDefines(std::ostream&) override {}
    void candidateDefines(VSpellCheck*) override {}
};


//======================================================================
// Text macros: the part of V3PreProc the test designs need (`define with and without arguments,
// `undef, `ifdef/`ifndef/`elsif/`else/`endif; VERILATOR, verilator, verilator3 and SYSTEMVERILOG are
// predefined).  Line structure is kept
// (a directive line becomes an empty line, continued defines leave empty lines) so that FileLines
// still point at the source.

class MacroExpander final {
    struct Macro final {
        bool hasArgs = false;
        std::vector<std::string> params;
        std::string body;
    };
    std::map<std::string, Macro> m_macros;
    std::string m_filename;  // for `__FILE__
    int m_line = 1;  // for `__LINE__

    static bool idStart(char c) { return std::isalpha(static_cast<unsigned char>(c)) || c == '_'; }
    static bool idChar(char c) { return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '$'; }
    static std::string trim(const std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
        while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
        return s.substr(a, b - a);
    }
    // expand the macro uses in one piece of text (no directives inside)
    std::string expand(const std::string& in, int depth = 0) const {
        if (depth > 40) return in;
        std::string out;
        for (size_t i = 0; i < in.size();) {
            const char c = in[i];
            if (c == '"') {  // string literal: copied verbatim
                size_t j = i + 1;
                while (j < in.size() && in[j] != '"') j += (in[j] == '\\') ? 2 : 1;
                out.append(in, i, std::min(j + 1, in.size()) - i);
                i = std::min(j + 1, in.size());
            } else if (c == '/' && i + 1 < in.size() && in[i + 1] == '/') {
                size_t j = in.find('\n', i);
                if (j == std::string::npos) j = in.size();
                i = j;  // line comment dropped (it may sit inside a macro body)
            } else if (c == '`' && i + 1 < in.size() && idStart(in[i + 1])) {
                size_t j = i + 1;
                while (j < in.size() && idChar(in[j])) ++j;
                const std::string name = in.substr(i + 1, j - i - 1);
                if (name == "__FILE__" || name == "__LINE__") {
                    out += name == "__FILE__" ? "\"" + m_filename + "\"" : std::to_string(m_line);
                    i = j;
                    continue;
                }
                const auto it = m_macros.find(name);
                if (it == m_macros.end()) {  // `line, `timescale ...: left for the lexer
                    out.append(in, i, j - i);
                    i = j;
                    continue;
                }
                const Macro& m = it->second;
                std::string body = m.body;
                if (m.hasArgs) {
                    size_t k = j;
                    while (k < in.size() && std::isspace(static_cast<unsigned char>(in[k]))) ++k;
                    if (k < in.size() && in[k] == '(') {
                        std::vector<std::string> args;
                        std::string cur;
                        int nest = 0;
                        size_t q = k + 1;
                        for (; q < in.size(); ++q) {
                            const char a = in[q];
                            if (a == '"') {
                                const size_t e = in.find('"', q + 1);
                                cur.append(in, q, (e == std::string::npos ? in.size() : e + 1) - q);
                                q = (e == std::string::npos) ? in.size() : e;
                                continue;
                            }
                            if (a == '(' || a == '[' || a == '{') ++nest;
                            if (a == ')' || a == ']' || a == '}') {
                                if (a == ')' && nest == 0) break;
                                --nest;
                            }
                            if (a == ',' && nest == 0) {
                                args.push_back(trim(cur));
                                cur.clear();
                            } else {
                                cur += a;
                            }
                        }
                        args.push_back(trim(cur));
                        j = std::min(q + 1, in.size());
                        // substitute parameters by name (identifier boundaries)
                        std::string sub;
                        for (size_t b = 0; b < body.size();) {
                            if (idStart(body[b])) {
                                size_t e = b;
                                while (e < body.size() && idChar(body[e])) ++e;
                                const std::string id = body.substr(b, e - b);
                                size_t pi = 0;
                                for (; pi < m.params.size(); ++pi) {
                                    if (m.params[pi] == id) break;
                                }
                                sub += (pi < m.params.size() && pi < args.size()) ? args[pi] : id;
                                b = e;
                            } else if (body[b] == '`' && b + 1 < body.size() && body[b + 1] == '"') {
                                sub += '"';  // `" : string quote inside a macro
                                b += 2;
                            } else if (body[b] == '`' && b + 1 < body.size() && body[b + 1] == '`') {
                                b += 2;  // `` : token paste
                            } else {
                                sub += body[b++];
                            }
                        }
                        body = sub;
                    }
                }
                out += expand(body, depth + 1);
                i = j;
            } else {
                out += c;
                ++i;
            }
        }
        return out;
    }

public:
    MacroExpander() {  // V3PreShell.cpp:53-60
        m_macros["VERILATOR"].body = "1";
        m_macros["verilator"].body = "1";
        m_macros["verilator3"].body = "1";
        m_macros["SYSTEMVERILOG"].body = "1";
    }
    std::string run(const std::string& src, const std::string& filename) {
        m_filename = filename;
        m_line = 1;
        std::string out;
        std::vector<std::pair<bool, bool>> conds;  // (this branch active, some branch taken)
        const auto active = [&]() {
            for (const auto& c : conds) {
                if (!c.first) return false;
            }
            return true;
        };
        size_t pos = 0;
        while (pos < src.size()) {
            size_t eol = src.find('\n', pos);
            if (eol == std::string::npos) eol = src.size();
            std::string line = src.substr(pos, eol - pos);
            int extraLines = 0;
            const std::string t = trim(line);
            const auto directive = [&](const char* name) {
                const size_t n = std::strlen(name);
                return t.compare(0, n, name) == 0 && (t.size() == n || !idChar(t[n]));
            };
            if (directive("`define")) {
                // continuation lines
                while (!line.empty() && line.back() == '\\' && eol < src.size()) {
                    line.pop_back();
                    const size_t next = src.find('\n', eol + 1);
                    const size_t e2 = (next == std::string::npos) ? src.size() : next;
                    line += ' ' + src.substr(eol + 1, e2 - eol - 1);
                    eol = e2;
                    ++extraLines;
                }
                if (active()) {
                    std::string rest = trim(trim(line).substr(7));
                    size_t k = 0;
                    while (k < rest.size() && idChar(rest[k])) ++k;
                    const std::string name = rest.substr(0, k);
                    Macro m;
                    if (k < rest.size() && rest[k] == '(') {
                        m.hasArgs = true;
                        const size_t close = rest.find(')', k);
                        std::string plist = rest.substr(k + 1, close - k - 1);
                        std::istringstream is{plist};
                        std::string p;
                        while (std::getline(is, p, ',')) {
                            p = trim(p);
                            const size_t eq = p.find('=');
                            if (eq != std::string::npos) p = trim(p.substr(0, eq));
                            m.params.push_back(p);
                        }
                        k = (close == std::string::npos) ? rest.size() : close + 1;
                    }
                    m.body = trim(rest.substr(k));
                    const size_t cmt = m.body.find("//");
                    if (cmt != std::string::npos && m.body.find('"') == std::string::npos) m.body = trim(m.body.substr(0, cmt));
                    m_macros[name] = m;
                }
                out += '\n';
            } else if (directive("`undef")) {
                if (active()) m_macros.erase(trim(t.substr(6)));
                out += '\n';
            } else if (directive("`ifdef") || directive("`ifndef")) {
                const bool neg = directive("`ifndef");
                std::string name = trim(t.substr(neg ? 7 : 6));
                size_t k = 0;
                while (k < name.size() && idChar(name[k])) ++k;
                name = name.substr(0, k);
                const bool on = (m_macros.count(name) != 0) != neg;
                conds.emplace_back(on, on);
                out += '\n';
            } else if (directive("`elsif")) {
                std::string name = trim(t.substr(6));
                size_t k = 0;
                while (k < name.size() && idChar(name[k])) ++k;
                name = name.substr(0, k);
                if (!conds.empty()) {
                    const bool on = !conds.back().second && m_macros.count(name) != 0;
                    conds.back().first = on;
                    conds.back().second = conds.back().second || on;
                }
                out += '\n';
            } else if (directive("`else")) {
                if (!conds.empty()) {
                    conds.back().first = !conds.back().second;
                    conds.back().second = true;
                }
                out += '\n';
            } else if (directive("`endif")) {
                if (!conds.empty()) conds.pop_back();
                out += '\n';
            } else {
                if (active()) out += expand(line);
                out += '\n';
            }
            for (int k = 0; k < extraLines; ++k) out += '\n';
            m_line += 1 + extraLines;
            pos = eol + 1;
        }
        return out;
    }
};

//======================================================================
// Lexer

enum class Tok : uint8_t { END, ID, SYSID, NUM, STR, PUNCT, PRAGMA };

struct Token final {
    Tok kind = Tok::END;
    std::string text;
    FileLine* fl = nullptr;
};

class Lexer final {
    const std::string m_src;
    size_t m_pos = 0;
    int m_line = 1;
    size_t m_lineStart = 0;
    std::string m_filename;
    FileLine* const m_baseFl;
    std::deque<Token> m_ahead;
    std::string m_pragma;  // "no_inline_module" / "inline_module" seen by skipSpace()
    size_t m_pragmaBegin = 0;

    FileLine* makeFl(size_t begin, size_t end) const {
        FileLine* const fl = new FileLine{m_baseFl};
        if (!m_filename.empty()) fl->filename(m_filename);
        fl->lineno(m_line);
        fl->column(static_cast<int>(begin - m_lineStart) + 1,
                   static_cast<int>(end - m_lineStart) + 1);
        return fl;
    }
    void newline() {
        ++m_line;
        m_lineStart = m_pos;
    }
    void skipSpace() {
        for (;;) {
            if (m_pos >= m_src.size()) return;
            const char c = m_src[m_pos];
            if (c == '\n') {
                ++m_pos;
                newline();
            } else if (std::isspace(static_cast<unsigned char>(c))) {
                ++m_pos;
            } else if (c == '/' && m_pos + 1 < m_src.size() && m_src[m_pos + 1] == '/') {
                while (m_pos < m_src.size() && m_src[m_pos] != '\n') ++m_pos;
            } else if (c == '/' && m_pos + 1 < m_src.size() && m_src[m_pos + 1] == '*') {
                // /*verilator no_inline_module*/ and /*verilator inline_module*/ (verilog.l:832-839;
                // V3PreProc::comment accepts surrounding blanks): the only meta-comments this
                // front end turns into tokens, all others stay comments
                {
                    const size_t end = m_src.find("*/", m_pos + 2);
                    if (end != std::string::npos) {
                        std::string text = m_src.substr(m_pos + 2, end - m_pos - 2);
                        while (!text.empty() && std::isspace(static_cast<unsigned char>(text.back()))) text.pop_back();
                        size_t lead = 0;
                        while (lead < text.size() && std::isspace(static_cast<unsigned char>(text[lead]))) ++lead;
                        text.erase(0, lead);
                        if (text == "verilator no_inline_module" || text == "verilator inline_module") {
                            m_pragma = text.substr(10);
                            m_pragmaBegin = m_pos;
                            m_pos = end + 2;
                            return;
                        }
                    }
                }
                m_pos += 2;
                while (m_pos + 1 < m_src.size() && !(m_src[m_pos] == '*' && m_src[m_pos + 1] == '/')) {
                    if (m_src[m_pos] == '\n') {
                        ++m_pos;
                        newline();
                    } else {
                        ++m_pos;
                    }
                }
                m_pos += 2;
            } else if (c == '`') {  // only the directives the preprocessor itself leaves behind
                const size_t eol = m_src.find('\n', m_pos);
                const std::string dir = m_src.substr(m_pos, eol - m_pos);
                if (dir.compare(0, 5, "`line") == 0) {
                    // `line N "file" level
                    std::istringstream is{dir.substr(5)};
                    int n = 1;
                    std::string fname;
                    is >> n >> fname;
                    if (fname.size() >= 2) m_filename = fname.substr(1, fname.size() - 2);
                    m_pos = (eol == std::string::npos) ? m_src.size() : eol + 1;
                    m_line = n;
                    m_lineStart = m_pos;
                } else {  // `begin_keywords, `timescale, `default_nettype ...: not interpreted
                    m_pos = (eol == std::string::npos) ? m_src.size() : eol;
                }
            } else {
                return;
            }
        }
    }
    static bool isIdStart(char c) { return std::isalpha(static_cast<unsigned char>(c)) || c == '_'; }
    static bool isIdChar(char c) {
        return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '$';
    }
    Token lexOne() {
        skipSpace();
        Token t;
        if (!m_pragma.empty()) {  // a meta-comment skipSpace() stopped behind
            t.kind = Tok::PRAGMA;
            t.text = m_pragma;
            t.fl = makeFl(m_pragmaBegin, m_pos);
            m_pragma.clear();
            return t;
        }
        if (m_pos >= m_src.size()) {
            t.fl = makeFl(m_pos, m_pos);
            return t;
        }
        const size_t begin = m_pos;
        const char c = m_src[m_pos];

        if (isIdStart(c)) {
            while (m_pos < m_src.size() && isIdChar(m_src[m_pos])) ++m_pos;
            t.kind = Tok::ID;
        } else if (c == '$' && m_pos + 1 < m_src.size() && isIdStart(m_src[m_pos + 1])) {
            ++m_pos;
            while (m_pos < m_src.size() && isIdChar(m_src[m_pos])) ++m_pos;
            t.kind = Tok::SYSID;
        } else if (std::isdigit(static_cast<unsigned char>(c)) || (c == '\'' && m_pos + 1 < m_src.size()
                                                                   && std::strchr("01xXzZsSbBoOdDhH", m_src[m_pos + 1]))) {
            // [size] ['[s]base digits]  |  'unbased  (verilog.l: integer literal patterns)
            while (m_pos < m_src.size() && (std::isdigit(static_cast<unsigned char>(m_src[m_pos])) || m_src[m_pos] == '_')) ++m_pos;
            size_t look = m_pos;
            while (look < m_src.size() && (m_src[look] == ' ' || m_src[look] == '\t')) ++look;
            if (look < m_src.size() && m_src[look] == '\'') {
                size_t p = look + 1;
                if (p < m_src.size() && (m_src[p] == 's' || m_src[p] == 'S')) ++p;
                if (p < m_src.size() && std::strchr("bBoOdDhH", m_src[p])) {
                    ++p;
                    while (p < m_src.size() && (m_src[p] == ' ' || m_src[p] == '\t')) ++p;
                    while (p < m_src.size() && (std::isxdigit(static_cast<unsigned char>(m_src[p]))
                                                || std::strchr("xXzZ?_", m_src[p]))) ++p;
                    m_pos = p;
                } else if (begin == look && p < m_src.size() && std::strchr("01xXzZ", m_src[p])) {
                    m_pos = p + 1;  // '0 '1 'x 'z
                }
            }
            t.kind = Tok::NUM;
        } else if (c == '"') {
            ++m_pos;
            while (m_pos < m_src.size() && m_src[m_pos] != '"') {
                if (m_src[m_pos] == '\\') ++m_pos;
                ++m_pos;
            }
            ++m_pos;
            t.kind = Tok::STR;
        } else {
            static const char* const puncts[] = {"<<<=", ">>>=", "===", "!==", "==?", "!=?", "<<<", ">>>", "<<=",
                                                 ">>=", "<->", "&&&", "**", "==", "!=", "<=", ">=", "&&", "||",
                                                 "<<", ">>", "+:", "-:", "->", "~&", "~|", "~^", "^~", "++", "--",
                                                 "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "::", ".*"};
            size_t len = 1;
            for (const char* p : puncts) {
                const size_t n = std::strlen(p);
                if (m_src.compare(m_pos, n, p) == 0) {
                    len = n;
                    break;
                }
            }
            m_pos += len;
            t.kind = Tok::PUNCT;
        }
        t.text = m_src.substr(begin, m_pos - begin);
        t.fl = makeFl(begin, m_pos);
        return t;
    }

public:
    Lexer(FileLine* baseFl, const std::string& src)
        : m_src{src}
        , m_baseFl{baseFl} {}
    const Token& peek(size_t n = 0) {
        while (m_ahead.size() <= n) m_ahead.push_back(lexOne());
        return m_ahead[n];
    }
    Token next() {
        peek();
        Token t = m_ahead.front();
        m_ahead.pop_front();
        return t;
    }
};

//======================================================================
// Parser

struct SyntaxError final {};

class Parser final {
    Lexer& m_lex;
    AstNetlist* const m_rootp;
    const bool m_inLibrary;
    const std::string m_libname;

    // Declaration state, the GRAMMARP->m_var* members of V3ParseGrammar.h:38-60
    VVarType m_varDecl = VVarType::UNKNOWN;
    VDirection m_varIO = VDirection::NONE;
    AstNodeDType* m_varDTypep = nullptr;
    bool m_varDeclTyped = false;
    bool m_pinAnsi = false;
    int m_pinNum = -1;
    // Instance being declared (INSTPREP, verilog.y:123-129)
    FileLine* m_instModuleFl = nullptr;
    std::string m_instModule;

    // -- token helpers ----------------------------------------------------
    [[noreturn]] void syntaxError(const Token& t, const std::string& what) {
        t.fl->v3error("V3ParseLite: syntax error, unexpected '" << t.text << "'"
                                                                << (what.empty() ? "" : ", expecting " + what));
        throw SyntaxError{};
    }
    bool isPunct(const char* p, size_t n = 0) {
        const Token& t = m_lex.peek(n);
        return t.kind == Tok::PUNCT && t.text == p;
    }
    bool isKw(const char* p, size_t n = 0) {
        const Token& t = m_lex.peek(n);
        return t.kind == Tok::ID && t.text == p;
    }
    bool acceptPunct(const char* p) {
        if (!isPunct(p)) return false;
        m_lex.next();
        return true;
    }
    bool acceptKw(const char* p) {
        if (!isKw(p)) return false;
        m_lex.next();
        return true;
    }
    Token expectPunct(const char* p) {
        if (!isPunct(p)) syntaxError(m_lex.peek(), std::string{"'"} + p + "'");
        return m_lex.next();
    }
    Token expectKw(const char* p) {
        if (!isKw(p)) syntaxError(m_lex.peek(), std::string{"'"} + p + "'");
        return m_lex.next();
    }
    Token expectId() {
        if (m_lex.peek().kind != Tok::ID || isReserved(m_lex.peek().text)) syntaxError(m_lex.peek(), "identifier");
        return m_lex.next();
    }
    static bool isReserved(const std::string& s) {
        static const std::set<std::string> kw = {
            "module", "endmodule", "input", "output", "inout", "wire", "reg", "logic", "bit", "integer", "int",
            "signed", "unsigned", "assign", "always", "always_ff", "always_comb", "always_latch", "initial",
            "begin", "end", "if", "else", "case", "casez", "casex", "endcase", "default", "for", "posedge",
            "negedge", "or", "parameter", "localparam", "assert", "cover", "property", "genvar", "generate",
            "endgenerate", "function", "endfunction", "task", "endtask", "final", "while", "do", "unique", "priority"};
        return kw.count(s) != 0;
    }

    // -- declaration helpers -------------------------------------------------
    // VARRESET_LIST / VARRESET_NONLIST (verilog.y:85-95)
    void varReset(VVarType decl, int pinNumStart) {
        m_varDecl = decl;
        m_varIO = VDirection::NONE;
        setDType(nullptr);
        m_pinNum = pinNumStart;
        m_varDeclTyped = false;
    }
    void setDType(AstNodeDType* dtypep) {  // V3ParseGrammar.h:165-169
        if (m_varDTypep) VL_DO_DANGLING(m_varDTypep->deleteTree(), m_varDTypep);
        m_varDTypep = dtypep;
    }
    bool tracingOn(const FileLine* fl) const { return v3Global.opt.trace() && fl->tracingOn(); }

    // createArray for unpacked ranges (V3ParseGrammar.cpp:131-173): innermost declared last
    static AstNodeDType* createUnpackArray(AstNodeDType* basep, std::vector<AstRange*>& ranges) {
        AstNodeDType* arrayp = basep;
        for (auto it = ranges.rbegin(); it != ranges.rend(); ++it) {
            arrayp = new AstUnpackArrayDType{(*it)->fileline(), VFlagChildDType{}, arrayp, *it};
        }
        return arrayp;
    }
    // addRange with one packed range (V3ParseGrammar.h:189-218): an implicit logic that gets a
    // range is no longer implicit
    static AstNodeDType* addRange(AstBasicDType* dtypep, AstRange* rangep) {
        if (!rangep) return dtypep;
        if (dtypep->implicit()) {
            AstBasicDType* const newp = new AstBasicDType{dtypep->fileline(), VBasicDTypeKwd::LOGIC,
                                                          dtypep->numeric(), dtypep->width(),
                                                          dtypep->widthMin()};
            VL_DO_DANGLING(dtypep->deleteTree(), dtypep);
            dtypep = newp;
        }
        dtypep->rangep(rangep);
        return dtypep;
    }
    // createVariable (V3ParseGrammar.cpp:175-266) for the declaration kinds this front end reads
    AstVar* createVariable(FileLine* fl, const std::string& name, std::vector<AstRange*>& arrayRanges) {
        AstNodeDType* dtypep = m_varDTypep ? m_varDTypep->cloneTree(false)
                                           : new AstBasicDType{fl, VBasicDTypeKwd::LOGIC_IMPLICIT};
        VVarType type = m_varDecl;
        if (type == VVarType::UNKNOWN) {
            UASSERT_OBJ(m_varIO.isAny(), fl, "Unknown signal type declared");
            type = VVarType::PORT;
        }
        AstNodeDType* const arrayDTypep = createUnpackArray(dtypep, arrayRanges);
        AstVar* const nodep = new AstVar{fl, type, name, VFlagChildDType{}, arrayDTypep};
        nodep->ansi(m_pinAnsi);
        nodep->declTyped(m_varDeclTyped);
        nodep->lifetime(VLifetime::NONE);
        if (m_varDecl != VVarType::UNKNOWN) nodep->combineType(m_varDecl);
        if (m_varIO != VDirection::NONE) {
            nodep->declDirection(m_varIO);
            nodep->direction(m_varIO);
        }
        if (nodep->isGenVar()) {
            nodep->trace(false);
        } else if (nodep->isParam() && !v3Global.opt.traceParams()) {
            nodep->trace(false);
        } else {
            nodep->trace(tracingOn(nodep->fileline()));
        }
        return nodep;
    }

    // '[' constExpr ':' constExpr ']'  (anyrange, verilog.y:3149-3151)
    AstRange* parseRange() {
        const Token lb = expectPunct("[");
        AstNodeExpr* const leftp = parseExpr();
        expectPunct(":");
        AstNodeExpr* const rightp = parseExpr();
        expectPunct("]");
        return new AstRange{lb.fl, leftp, rightp};
    }
    // variable_dimension: anyrange, or '[' constExpr ']' = [0:size-1] (verilog.y:2303-2309)
    AstRange* parseUnpackedDim() {
        const Token lb = expectPunct("[");
        AstNodeExpr* const leftp = parseExpr();
        if (acceptPunct(":")) {
            AstNodeExpr* const rightp = parseExpr();
            expectPunct("]");
            return new AstRange{lb.fl, leftp, rightp};
        }
        expectPunct("]");
        return new AstRange{lb.fl, new AstConst{lb.fl, 0}, new AstSub{lb.fl, leftp, new AstConst{lb.fl, 1}}};
    }
    VSigning parseSigningE() {
        if (acceptKw("signed")) return VSigning::SIGNED;
        if (acceptKw("unsigned")) return VSigning::UNSIGNED;
        return VSigning::NOSIGN;
    }
    // data_type = integer_vector_type signingE rangeListE | integer_atom_type signingE
    // (data_typeBasic, verilog.y:2096-2100); returns nullptr when the next token is no type keyword
    AstNodeDType* parseDataTypeE() {
        const Token& t = m_lex.peek();
        if (t.kind != Tok::ID) return nullptr;
        if (t.text == "reg" || t.text == "logic" || t.text == "bit") {
            const Token kw = m_lex.next();
            AstBasicDType* const bp = new AstBasicDType{
                kw.fl, kw.text == "bit" ? VBasicDTypeKwd::BIT : VBasicDTypeKwd::LOGIC};
            bp->setSignedState(parseSigningE());
            return addRange(bp, isPunct("[") ? parseRange() : nullptr);
        }
        if (t.text == "integer" || t.text == "int") {
            const Token kw = m_lex.next();
            AstBasicDType* const bp = new AstBasicDType{
                kw.fl, kw.text == "int" ? VBasicDTypeKwd::INT : VBasicDTypeKwd::INTEGER};
            bp->setSignedState(parseSigningE());
            return bp;
        }
        return nullptr;
    }

    // -- module ---------------------------------------------------------------
    // module_declaration / modFront (verilog.y:1200-1236)
    void parseModule() {
        expectKw("module");
        const Token name = expectId();
        AstModule* const modp = new AstModule{name.fl, name.text, m_libname};
        modp->inLibrary(m_inLibrary || modp->fileline()->celldefineOn());
        modp->modTrace(tracingOn(modp->fileline()));
        modp->timeunit(VTimescale{VTimescale::NONE});
        m_rootp->timeprecisionMerge(modp->fileline(), VTimescale{VTimescale::NONE});
        m_rootp->addModulesp(modp);
        bool hasParamList = false;
        if (isPunct("#")) {  // parameter_port_list (verilog.y:1261-1300)
            m_lex.next();
            expectPunct("(");
            hasParamList = true;
            m_pinAnsi = true;
            varReset(VVarType::GPARAM, 1);
            while (!isPunct(")")) {
                if (acceptKw("parameter")) {
                    varReset(VVarType::GPARAM, m_pinNum);
                } else if (acceptKw("localparam")) {
                    varReset(VVarType::LPARAM, m_pinNum);
                }
                parseParamTypeE();
                modp->addStmtsp(parseParamAssignment());
                if (!acceptPunct(",")) break;
            }
            expectPunct(")");
            m_pinAnsi = false;
            varReset(VVarType::UNKNOWN, 0);
        }
        modp->hasParameterList(hasParamList);
        if (acceptPunct("(")) {  // portsStarE: '(' list_of_ports ')' (verilog.y:1329-1332)
            varReset(VVarType::PORT, 1);
            // list_of_ports of plain identifiers: the Verilog-1995 form, directions and types
            // follow as port_declaration module items (verilog.y:1424-1437, 1540-1575)
            const bool nonAnsi = m_lex.peek().kind == Tok::ID && !isKw("input") && !isKw("output")
                                 && !isKw("inout") && !isKw("wire") && !isKw("reg") && !isKw("logic")
                                 && !isKw("bit") && !isKw("integer") && !isKw("int") && !isKw("signed")
                                 && !isKw("unsigned");
            m_pinAnsi = !nonAnsi;
            if (nonAnsi) {
                do {
                    const Token name = expectId();
                    modp->addStmtsp(new AstPort{name.fl, m_pinNum++, name.text});
                } while (acceptPunct(","));
            } else if (!isPunct(")")) {
                do { parsePort(modp); } while (acceptPunct(","));
            }
            expectPunct(")");
            varReset(VVarType::UNKNOWN, 0);
            m_pinAnsi = false;
        }
        expectPunct(";");
        while (!isKw("endmodule")) {
            if (m_lex.peek().kind == Tok::END) syntaxError(m_lex.peek(), "'endmodule'");
            if (AstNode* const itemp = parseModuleItem()) modp->addStmtsp(itemp);
        }
        expectKw("endmodule");
        if (acceptPunct(":")) expectId();
        setDType(nullptr);
    }

    // port (ANSI forms, verilog.y:1438-1480) with portDirNetE (:1491-1498)
    void parsePort(AstModule* modp) {
        bool gotDir = false;
        if (isKw("input") || isKw("output") || isKw("inout")) {
            const Token dir = m_lex.next();
            m_varIO = dir.text == "input"    ? VDirection::INPUT
                      : dir.text == "output" ? VDirection::OUTPUT
                                             : VDirection::INOUT;
            m_varDecl = VVarType::PORT;
            gotDir = true;
        }
        bool gotNet = false;
        if (acceptKw("wire")) {
            m_varDecl = VVarType::WIRE;
            gotNet = true;
        }
        if (gotDir || gotNet) {
            setDType(nullptr);  // VARDTYPE_NDECL(nullptr)
            m_varDeclTyped = false;
        }
        if (AstNodeDType* const dtp = parseDataTypeE()) {  // portDirNetE data_type portSig
            setDType(dtp);
            m_varDeclTyped = true;
            if (m_varIO == VDirection::NONE) m_varIO = VDirection::INPUT;  // VARIOANSI
        } else if (isKw("signed") || isKw("unsigned") || isPunct("[")) {
            // portDirNetE signingE rangeList portSig / portDirNetE signing portSig
            const VSigning sign = parseSigningE();
            AstRange* const rangep = isPunct("[") ? parseRange() : nullptr;
            FileLine* const dfl = rangep ? rangep->fileline() : m_lex.peek().fl;
            setDType(addRange(new AstBasicDType{dfl, VBasicDTypeKwd::LOGIC_IMPLICIT, sign}, rangep));
            if (!rangep && m_varIO == VDirection::NONE) m_varIO = VDirection::INPUT;
        }
        const Token name = expectId();  // portSig
        AstPort* const portp = new AstPort{name.fl, m_pinNum++, name.text};
        modp->addStmtsp(portp);
        std::vector<AstRange*> dims;
        while (isPunct("[")) dims.push_back(parseUnpackedDim());
        AstVar* const varp = createVariable(name.fl, name.text, dims);
        modp->addStmtsp(varp);
        if (acceptPunct("=")) varp->valuep(parseExpr());
    }

    // genItemOrBegin / genItemBegin / generate_block_or_null / loop_generate_construct /
    // conditional_generate_construct (if) (verilog.y:2836-2953)
    AstNode* parseGenItemOrBegin() {
        if (isKw("begin") || (m_lex.peek().kind == Tok::ID && !isReserved(m_lex.peek().text) && m_lex.peek(1).kind == Tok::PUNCT
                              && m_lex.peek(1).text == ":" && m_lex.peek(2).kind == Tok::ID && m_lex.peek(2).text == "begin")) {
            return parseGenBegin();
        }
        if (isKw("for")) {
            const Token kw = m_lex.next();
            expectPunct("(");
            AstNode* varp = nullptr;
            if (acceptKw("genvar")) {  // yGENVAR genvar_identifierDecl '=' constExpr
                varReset(VVarType::GENVAR, 0);
                setDType(new AstBasicDType{kw.fl, VBasicDTypeKwd::INTEGER});
                m_varDeclTyped = true;
                const Token name = expectId();
                std::vector<AstRange*> dims;
                varp = createVariable(name.fl, name.text, dims);
                const Token eq = expectPunct("=");
                varp->addNext(new AstAssign{eq.fl, new AstParseRef{name.fl, name.text, nullptr, nullptr}, parseExpr()});
                varReset(VVarType::UNKNOWN, 0);
            }
            AstNode* initp = varp;
            if (!varp) {
                const Token name = expectId();
                const Token eq = expectPunct("=");
                initp = new AstAssign{eq.fl, new AstParseRef{name.fl, name.text, nullptr, nullptr}, parseExpr()};
            }
            expectPunct(";");
            AstNodeExpr* const condp = parseExpr();
            expectPunct(";");
            AstNode* const incp = parseGenvarIteration();
            expectPunct(")");
            AstNode* const bodyp = parseGenBlockOrNull();
            // BEGIN(GENFOR(...)): the begin hides the local genvar (verilog.y:2929-2951)
            AstGenBlock* const lowerp = VN_CAST(bodyp, GenBlock);
            AstNode* const itemsp = lowerp && lowerp->itemsp() ? lowerp->itemsp()->unlinkFrBackWithNext() : nullptr;
            AstGenBlock* const blkp = new AstGenBlock{kw.fl, lowerp ? lowerp->name() : "", nullptr, true};
            if (VN_IS(initp, Var)) {
                AstNode* const assignp = initp->nextp();
                assignp->unlinkFrBackWithNext();
                blkp->addItemsp(initp);
                initp = assignp;
            }
            blkp->genforp(new AstGenFor{kw.fl, initp, condp, incp, itemsp});
            if (lowerp) VL_DO_DANGLING(lowerp->deleteTree(), lowerp);
            return blkp;
        }
        if (isKw("if")) {
            const Token kw = m_lex.next();
            expectPunct("(");
            AstNodeExpr* const condp = parseExpr();
            expectPunct(")");
            AstNode* const thenp = parseGenBlockOrNull();
            AstNode* elsep = nullptr;
            if (acceptKw("else")) elsep = parseGenBlockOrNull();
            return new AstGenIf{kw.fl, condp, thenp, elsep};
        }
        return parseModuleItem();
    }
    AstNode* parseGenBegin() {
        std::string name;
        FileLine* fl = m_lex.peek().fl;
        if (!isKw("begin")) {  // id ':' begin
            name = m_lex.next().text;
            m_lex.next();
        }
        const Token kw = m_lex.next();  // begin
        if (acceptPunct(":")) {
            const Token lbl = expectId();
            name = lbl.text;
            fl = lbl.fl;
        }
        AstNode* listp = nullptr;
        while (!isKw("end")) {
            if (m_lex.peek().kind == Tok::END) syntaxError(m_lex.peek(), "'end'");
            listp = AstNode::addNextNull(listp, parseGenItemOrBegin());
        }
        m_lex.next();
        if (acceptPunct(":")) expectId();
        if (!listp && name.empty()) return nullptr;
        return new AstGenBlock{name.empty() ? kw.fl : fl, name, listp, false};
    }
    AstNode* parseGenBlockOrNull() {
        if (isKw("begin") || (m_lex.peek().kind == Tok::ID && m_lex.peek(1).kind == Tok::PUNCT && m_lex.peek(1).text == ":"
                              && m_lex.peek(2).kind == Tok::ID && m_lex.peek(2).text == "begin")) {
            return parseGenBegin();
        }
        AstNode* const itemp = parseGenItemOrBegin();
        return itemp ? new AstGenBlock{itemp->fileline(), "", itemp, true} : nullptr;
    }
    AstNode* parseGenvarIteration() {  // genvar_iteration (verilog.y:2962-2999), ++/-- included
        const Token name = expectId();
        AstParseRef* const refp = new AstParseRef{name.fl, name.text, nullptr, nullptr};
        const Token op = m_lex.next();
        const auto ref2 = [&]() { return new AstParseRef{name.fl, name.text, nullptr, nullptr}; };
        if (op.text == "=") return new AstAssign{op.fl, refp, parseExpr()};
        if (op.text == "+=") return new AstAssign{op.fl, refp, new AstAdd{op.fl, ref2(), parseExpr()}};
        if (op.text == "-=") return new AstAssign{op.fl, refp, new AstSub{op.fl, ref2(), parseExpr()}};
        if (op.text == "++") return new AstAssign{op.fl, refp, new AstAdd{op.fl, ref2(), new AstConst{op.fl, AstConst::StringToParse{}, "'b1"}}};
        if (op.text == "--") return new AstAssign{op.fl, refp, new AstSub{op.fl, ref2(), new AstConst{op.fl, AstConst::StringToParse{}, "'b1"}}};
        syntaxError(op, "genvar iteration");
    }

    // function_declaration / task_declaration (verilog.y:4637-4660, 4702-4749, 4788-4830):
    // [automatic] funcId tfGuts: ANSI or Verilog-1995 ports, block item declarations, statements
    AstNode* parseFTask() {
        const Token kw = m_lex.next();
        const bool isFunc = kw.text == "function";
        VLifetime lifetime = VLifetime::NONE;
        if (m_lex.peek().kind == Tok::ID && m_lex.peek().text == "automatic") {
            m_lex.next();
            lifetime = VLifetime::AUTOMATIC_EXPLICIT;
        } else if (m_lex.peek().kind == Tok::ID && m_lex.peek().text == "static") {
            m_lex.next();
            lifetime = VLifetime::STATIC_EXPLICIT;
        }
        AstNodeDType* fvarp = nullptr;
        if (isFunc) {  // funcId: function_data_type_or_implicit
            if (AstNodeDType* const dtp = parseDataTypeE()) {
                fvarp = dtp;
            } else {
                const VSigning sign = parseSigningE();
                AstRange* const rangep = isPunct("[") ? parseRange() : nullptr;
                FileLine* const dfl = rangep ? rangep->fileline() : m_lex.peek().fl;
                fvarp = addRange(new AstBasicDType{dfl, VBasicDTypeKwd::LOGIC_IMPLICIT, sign}, rangep);
            }
        }
        const Token name = expectId();
        AstNodeFTask* ftaskp;
        if (isFunc) {
            ftaskp = new AstFunc{name.fl, name.text, nullptr, fvarp};
        } else {
            AstTask* const taskp = new AstTask{name.fl, name.text, nullptr};
            taskp->verilogTask(true);
            ftaskp = taskp;
        }
        ftaskp->lifetime(lifetime);
        // saved declaration state of the enclosing module
        const VVarType saveDecl = m_varDecl;
        const VDirection saveIO = m_varIO;
        const bool saveAnsi = m_pinAnsi;
        const int savePin = m_pinNum;
        const auto portDecl = [&](bool list) {  // direction [type] names
            const Token dir = m_lex.next();
            varReset(VVarType::UNKNOWN, m_pinNum);
            m_varIO = dir.text == "input" ? VDirection::INPUT : dir.text == "output" ? VDirection::OUTPUT : VDirection::INOUT;
            if (AstNodeDType* const dtp = parseDataTypeE()) {
                setDType(dtp);
                m_varDeclTyped = true;
            } else if (isKw("signed") || isKw("unsigned") || isPunct("[")) {
                const VSigning sign = parseSigningE();
                AstRange* const rangep = isPunct("[") ? parseRange() : nullptr;
                FileLine* const dfl = rangep ? rangep->fileline() : m_lex.peek().fl;
                setDType(addRange(new AstBasicDType{dfl, VBasicDTypeKwd::LOGIC_IMPLICIT, sign}, rangep));
            }
            (void)list;
        };
        if (acceptPunct("(")) {  // tf_port_listE
            varReset(VVarType::UNKNOWN, 0);
            m_varIO = VDirection::INPUT;
            while (!isPunct(")")) {
                if (isKw("input") || isKw("output") || isKw("inout")) portDecl(true);
                const Token pname = expectId();
                std::vector<AstRange*> dims;
                while (isPunct("[")) dims.push_back(parseUnpackedDim());
                AstVar* const varp = createVariable(pname.fl, pname.text, dims);
                varp->funcLocal(true);
                ftaskp->addStmtsp(varp);
                if (!acceptPunct(",")) break;
            }
            expectPunct(")");
        }
        expectPunct(";");
        const char* const endKw = isFunc ? "endfunction" : "endtask";
        while (!isKw(endKw)) {
            if (m_lex.peek().kind == Tok::END) syntaxError(m_lex.peek(), endKw);
            if (isKw("input") || isKw("output") || isKw("inout")) {  // tf_port_declaration
                portDecl(false);
                AstNode* const listp = parseDeclList(false);
                for (AstNode* np = listp; np; np = np->nextp()) {
                    if (AstVar* const varp = VN_CAST(np, Var)) varp->funcLocal(true);
                }
                ftaskp->addStmtsp(listp);
            } else if (isKw("reg") || isKw("logic") || isKw("integer") || isKw("int") || isKw("bit")) {
                varReset(VVarType::VAR, 0);  // block_item_declaration
                setDType(parseDataTypeE());
                m_varDeclTyped = true;
                AstNode* const listp = parseDeclList(false);
                for (AstNode* np = listp; np; np = np->nextp()) {
                    if (AstVar* const varp = VN_CAST(np, Var)) varp->funcLocal(true);
                }
                ftaskp->addStmtsp(listp);
            } else if (AstNodeStmt* const stmtp = parseStmt()) {
                ftaskp->addStmtsp(stmtp);
            }
        }
        m_lex.next();
        if (acceptPunct(":")) expectId();
        varReset(saveDecl, savePin);
        m_varIO = saveIO;
        m_pinAnsi = saveAnsi;
        return ftaskp;
    }

    // implicit_typeE | data_type after parameter/localparam (parameter_declarationFront :1796-1799)
    void parseParamTypeE() {
        if (AstNodeDType* const dtp = parseDataTypeE()) {
            setDType(dtp);
        } else if (isKw("signed") || isKw("unsigned") || isPunct("[")) {
            const VSigning sign = parseSigningE();
            AstRange* const rangep = isPunct("[") ? parseRange() : nullptr;
            FileLine* const dfl = rangep ? rangep->fileline() : m_lex.peek().fl;
            setDType(addRange(new AstBasicDType{dfl, VBasicDTypeKwd::LOGIC_IMPLICIT, sign}, rangep));
        } else {
            setDType(nullptr);
        }
        m_varDeclTyped = true;
    }
    // param_assignment (verilog.y:3188-3197)
    AstVar* parseParamAssignment() {
        const Token name = expectId();
        std::vector<AstRange*> dims;
        AstVar* const varp = createVariable(name.fl, name.text, dims);
        if (acceptPunct("=")) varp->valuep(parseExpr());
        return varp;
    }

    // module_item subset (verilog.y:2600-2800)
    AstNode* parseModuleItem() {
        const Token& t = m_lex.peek();
        if (t.kind == Tok::PUNCT && t.text == ";") {
            m_lex.next();
            return nullptr;
        }
        if (t.kind == Tok::PRAGMA) {  // verilog.y:2727-2728 (yVL_INLINE_MODULE / yVL_NO_INLINE_MODULE)
            const Token kw = m_lex.next();
            return new AstPragma{kw.fl, kw.text == "inline_module" ? VPragmaType::INLINE_MODULE
                                                                   : VPragmaType::NO_INLINE_MODULE};
        }
        if (t.kind != Tok::ID) syntaxError(t, "module item");
        if (t.text == "parameter" || t.text == "localparam") {
            const Token kw = m_lex.next();  // varParamReset (verilog.y:1920-1923)
            varReset(kw.text == "parameter" ? VVarType::GPARAM : VVarType::LPARAM, 0);
            parseParamTypeE();
            AstNode* listp = nullptr;
            do { listp = AstNode::addNext(listp, static_cast<AstNode*>(parseParamAssignment())); } while (acceptPunct(","));
            expectPunct(";");
            return listp;
        }
        if (t.text == "input" || t.text == "output" || t.text == "inout") {
            // port_declaration (verilog.y:1540-1575): port_directionReset port_declNetE
            // [data_type | signingE rangeList] list_of_variable_decl_assignments
            const Token dir = m_lex.next();
            varReset(VVarType::UNKNOWN, m_pinNum);
            m_varIO = dir.text == "input"    ? VDirection::INPUT
                      : dir.text == "output" ? VDirection::OUTPUT
                                             : VDirection::INOUT;
            if (acceptKw("wire")) m_varDecl = VVarType::WIRE;
            if (AstNodeDType* const dtp = parseDataTypeE()) {
                setDType(dtp);
                m_varDeclTyped = true;
            } else if (isKw("signed") || isKw("unsigned") || isPunct("[")) {
                const VSigning sign = parseSigningE();
                AstRange* const rangep = isPunct("[") ? parseRange() : nullptr;
                FileLine* const dfl = rangep ? rangep->fileline() : m_lex.peek().fl;
                setDType(addRange(new AstBasicDType{dfl, VBasicDTypeKwd::LOGIC_IMPLICIT, sign}, rangep));
            }
            AstNode* const listp = parseDeclList(false);
            varReset(VVarType::UNKNOWN, m_pinNum);
            return listp;
        }
        if (t.text == "function" || t.text == "task") return parseFTask();
        if (t.text == "generate") {  // generate_region (verilog.y:2823-2826)
            m_lex.next();
            AstNode* listp = nullptr;
            while (!isKw("endgenerate")) {
                if (m_lex.peek().kind == Tok::END) syntaxError(m_lex.peek(), "'endgenerate'");
                listp = AstNode::addNextNull(listp, parseGenItemOrBegin());
            }
            m_lex.next();
            return listp;
        }
        // loop_generate_construct / conditional_generate_construct outside a generate region
        if (t.text == "for" || t.text == "if" || t.text == "begin") return parseGenItemOrBegin();
        if (t.text == "wire") {  // net_declaration (verilog.y:1856-1875, 1888-1902)
            m_lex.next();
            varReset(VVarType::UNKNOWN, 0);
            m_varDecl = VVarType::WIRE;
            if (AstNodeDType* const dtp = parseDataTypeE()) {  // var_data_type
                setDType(dtp);
            } else {
                const VSigning sign = parseSigningE();
                if (isPunct("[")) {
                    AstRange* const rangep = parseRange();
                    setDType(addRange(new AstBasicDType{rangep->fileline(), VBasicDTypeKwd::LOGIC, sign}, rangep));
                } else {
                    setDType(new AstBasicDType{m_lex.peek().fl, VBasicDTypeKwd::LOGIC, sign});
                }
            }
            return parseDeclList(true);
        }
        if (t.text == "reg" || t.text == "logic" || t.text == "bit" || t.text == "integer" || t.text == "int") {
            // data_declarationVar (verilog.y:2431-2480): VARRESET_NONLIST(VAR); VARDTYPE(data_type)
            varReset(VVarType::VAR, 0);
            setDType(parseDataTypeE());
            m_varDeclTyped = true;
            return parseDeclList(false);
        }
        if (t.text == "genvar") {
            m_lex.next();
            varReset(VVarType::GENVAR, 0);
            setDType(new AstBasicDType{t.fl, VBasicDTypeKwd::INTEGER});
            m_varDeclTyped = true;
            return parseDeclList(false);
        }
        if (t.text == "assign") {  // continuous_assign / assignOne (verilog.y:2742-2747, 3029-3032)
            m_lex.next();
            AstNode* listp = nullptr;
            do {
                AstNodeExpr* const lhsp = parseLvalue();
                const Token eq = expectPunct("=");
                AstNodeExpr* const rhsp = parseExpr();
                AstAssignW* const ap = new AstAssignW{eq.fl, lhsp, rhsp};
                listp = AstNode::addNext(listp, static_cast<AstNode*>(new AstAlways{ap}));
            } while (acceptPunct(","));
            expectPunct(";");
            return listp;
        }
        if (t.text == "always" || t.text == "always_ff" || t.text == "always_comb" || t.text == "always_latch") {
            const Token kw = m_lex.next();  // always_construct (verilog.y:2735-2740)
            const VAlwaysKwd akw = kw.text == "always"        ? VAlwaysKwd::ALWAYS
                                   : kw.text == "always_ff"   ? VAlwaysKwd::ALWAYS_FF
                                   : kw.text == "always_comb" ? VAlwaysKwd::ALWAYS_COMB
                                                              : VAlwaysKwd::ALWAYS_LATCH;
            return new AstAlways{kw.fl, akw, nullptr, parseStmt()};
        }
        if (t.text == "initial") {
            const Token kw = m_lex.next();
            return new AstInitial{kw.fl, parseStmt()};
        }
        if (t.text == "final") {
            const Token kw = m_lex.next();
            return new AstFinal{kw.fl, parseStmt()};
        }
        if (t.text == "cover" || t.text == "assert" || (m_lex.peek(1).kind == Tok::PUNCT && m_lex.peek(1).text == ":"
                                                       && (isKw("cover", 2) || isKw("assert", 2)))) {
            // concurrent_assertion_item (verilog.y:6480-6485)
            if (isPunct(":", 1)) {
                const Token label = m_lex.next();
                m_lex.next();
                return new AstBegin{label.fl, label.text, parseConcurrentAssertion(), true};
            }
            return parseConcurrentAssertion();
        }
        if (!isReserved(t.text) && m_lex.peek(1).kind == Tok::ID && !isReserved(m_lex.peek(1).text)) {
            return parseInstance();
        }
        if (!isReserved(t.text) && isPunct("#", 1)) return parseInstance();
        syntaxError(t, "module item");
    }

    // list_of_variable_decl_assignments / netSigList (verilog.y:2252-2262, 3078-3092)
    AstNode* parseDeclList(bool isNet) {
        AstNode* listp = nullptr;
        do {
            const Token name = expectId();
            std::vector<AstRange*> dims;
            while (isPunct("[")) dims.push_back(parseUnpackedDim());
            AstVar* const varp = createVariable(name.fl, name.text, dims);
            listp = AstNode::addNext(listp, static_cast<AstNode*>(varp));
            if (isPunct("=")) {
                const Token eq = m_lex.next();
                AstNodeExpr* const valuep = parseExpr();
                if (isNet) {  // net_decl_assignment: an AssignW in an Always
                    AstAssignW* const ap
                        = new AstAssignW{eq.fl, new AstParseRef{name.fl, name.text}, valuep, nullptr};
                    listp = AstNode::addNext(listp, static_cast<AstNode*>(new AstAlways{ap}));
                } else {
                    varp->valuep(valuep);
                }
            }
        } while (acceptPunct(","));
        expectPunct(";");
        return listp;
    }

    // instDecl / instnameParen / instPinItemE (verilog.y:3274-3312, 3399-3424)
    AstNode* parseInstance() {
        const Token mod = expectId();
        m_instModuleFl = mod.fl;
        m_instModule = mod.text;
        AstPin* paramsp = nullptr;
        if (acceptPunct("#")) {
            expectPunct("(");
            paramsp = parsePinList();
            expectPunct(")");
        }
        AstNode* listp = nullptr;
        do {
            const Token inst = expectId();
            expectPunct("(");
            AstPin* const pinsp = parsePinList();
            expectPunct(")");
            AstCell* const cellp = new AstCell{inst.fl, m_instModuleFl, inst.text, m_instModule, pinsp,
                                               paramsp ? paramsp->cloneTree(true) : nullptr, nullptr};
            cellp->trace(tracingOn(inst.fl));
            listp = AstNode::addNext(listp, static_cast<AstNode*>(cellp));
        } while (acceptPunct(","));
        expectPunct(";");
        if (paramsp) VL_DO_DANGLING(paramsp->deleteTree(), paramsp);
        return listp;
    }
    AstPin* parsePinList() {
        AstPin* listp = nullptr;
        int pinNum = 1;
        if (isPunct(")")) return new AstPin{m_lex.peek().fl, pinNum, "", nullptr};
        do {
            AstPin* pinp;
            if (acceptPunct(".")) {
                const Token name = expectId();
                if (acceptPunct("(")) {
                    AstNodeExpr* const exprp = isPunct(")") ? nullptr : parseExpr();
                    expectPunct(")");
                    pinp = new AstPin{name.fl, pinNum++, name.text, exprp};
                    pinp->svDotName(true);
                } else {
                    pinp = new AstPin{name.fl, pinNum++, name.text,
                                      new AstParseRef{name.fl, name.text, nullptr, nullptr}};
                    pinp->svDotName(true);
                    pinp->svImplicit(true);
                }
            } else {
                AstNodeExpr* const exprp = parseExpr();
                pinp = new AstPin{exprp->fileline(), pinNum++, "", exprp};
            }
            if (listp) { listp->addNext(pinp); } else { listp = pinp; }
        } while (acceptPunct(","));
        return listp;
    }

    // concurrent_assertion_statement / property_spec (verilog.y:6488-6497, 6686-6704)
    AstNodeStmt* parseConcurrentAssertion() {
        const Token kw = m_lex.next();  // assert | cover
        expectKw("property");
        expectPunct("(");
        AstSenItem* sensesp = nullptr;
        FileLine* specFl = nullptr;
        if (isPunct("@")) {
            const Token at = m_lex.next();
            specFl = at.fl;
            expectPunct("(");
            sensesp = parseSenItem();
            expectPunct(")");
        }
        AstNodeExpr* const exprp = parseExpr();
        AstPropSpec* const specp = new AstPropSpec{exprp->fileline(), nullptr, nullptr, exprp};
        if (specFl) {
            specp->fileline(specFl);
            specp->sensesp(sensesp);
        }
        expectPunct(")");
        AstNode* const stmtp = parseStmt();
        if (kw.text == "cover") return new AstCover{kw.fl, specp, stmtp, VAssertType::CONCURRENT};
        return new AstAssert{kw.fl, specp, stmtp, nullptr, VAssertType::CONCURRENT,
                             VAssertDirectiveType::ASSERT};
    }

    // senitem / senitemEdge (verilog.y:3468-3498)
    AstSenItem* parseSenItem() {
        if (isKw("posedge") || isKw("negedge")) {
            const Token edge = m_lex.next();
            AstNodeExpr* const exprp = parseExpr();
            return new AstSenItem{edge.fl, edge.text == "posedge" ? VEdgeType::ET_POSEDGE : VEdgeType::ET_NEGEDGE,
                                  exprp};
        }
        FileLine* const fl = m_lex.peek().fl;
        return new AstSenItem{fl, VEdgeType::ET_CHANGED, parseExpr()};
    }
    // event_control (verilog.y:3440-3446)
    AstSenTree* parseEventControl() {
        const Token at = expectPunct("@");
        if (acceptPunct("*")) {  // createSenTreeDotStar (V3ParseGrammar.h:127-129)
            return new AstSenTree{at.fl, new AstSenItem{at.fl, VEdgeType::ET_COMBO_STAR, nullptr}};
        }
        expectPunct("(");
        if (acceptPunct("*")) {
            expectPunct(")");
            return new AstSenTree{at.fl, new AstSenItem{at.fl, VEdgeType::ET_COMBO_STAR, nullptr}};
        }
        AstSenItem* itemsp = parseSenItem();
        while (acceptKw("or") || acceptPunct(",")) itemsp->addNext(parseSenItem());
        expectPunct(")");
        return new AstSenTree{at.fl, itemsp};
    }

    // -- statements -------------------------------------------------------------
    // PARSEP->newBlock (V3ParseImp.h:264-267)
    static AstNode* newBlock(FileLine* fl, AstNode* nodep) {
        if (nodep) return nodep;
        return new AstBegin{fl, "", nullptr, true};
    }
    // stmt (verilog.y:3573-3579) + statement_item (:3584-3790)
    AstNodeStmt* parseStmt() {
        const Token& t = m_lex.peek();
        if (t.kind == Tok::PUNCT) {
            if (t.text == ";") {
                m_lex.next();
                return nullptr;
            }
            if (t.text == "@") {  // event_control stmt
                AstSenTree* const sensesp = parseEventControl();
                AstNodeStmt* const stmtp = parseStmt();
                return new AstEventControl{sensesp->fileline(), sensesp, stmtp};
            }
            if (t.text == "{") return parseAssignStmt();
            syntaxError(t, "statement");
        }
        if (t.kind == Tok::SYSID) {
            AstNodeStmt* const stmtp = parseSystemTask();
            expectPunct(";");
            return stmtp;
        }
        if (t.kind != Tok::ID) syntaxError(t, "statement");
        if (t.text == "begin") {  // seq_block (verilog.y:3501-3510)
            const Token kw = m_lex.next();
            std::string name;
            if (acceptPunct(":")) name = expectId().text;
            AstBegin* const beginp = new AstBegin{kw.fl, name, nullptr, false};
            while (!isKw("end")) {
                if (m_lex.peek().kind == Tok::END) syntaxError(m_lex.peek(), "'end'");
                if (isKw("reg") || isKw("logic") || isKw("integer") || isKw("int") || isKw("bit")) {
                    varReset(VVarType::VAR, 0);  // block_item_declaration
                    setDType(parseDataTypeE());
                    m_varDeclTyped = true;
                    beginp->addDeclsp(parseDeclList(false));
                } else if (AstNodeStmt* const stmtp = parseStmt()) {
                    beginp->addStmtsp(stmtp);
                }
            }
            expectKw("end");
            if (acceptPunct(":")) expectId();
            return beginp;
        }
        if (t.text == "if" || ((t.text == "unique" || t.text == "priority") && isKw("if", 1))) {
            bool uniq = false, prio = false;
            if (acceptKw("unique")) uniq = true;
            if (acceptKw("priority")) prio = true;
            const Token kw = m_lex.next();
            expectPunct("(");
            AstNodeExpr* const condp = parseExpr();
            expectPunct(")");
            AstNode* const thensp = newBlock(kw.fl, parseStmt());
            AstIf* ifp;
            if (acceptKw("else")) {
                ifp = new AstIf{kw.fl, condp, thensp, newBlock(kw.fl, parseStmt())};
            } else {
                ifp = new AstIf{kw.fl, condp, thensp};
            }
            if (uniq) ifp->uniquePragma(true);
            if (prio) ifp->priorityPragma(true);
            return ifp;
        }
        if (t.text == "case" || t.text == "casez" || t.text == "casex"
            || ((t.text == "unique" || t.text == "priority") && (isKw("case", 1) || isKw("casez", 1)))) {
            bool uniq = false, prio = false;
            if (acceptKw("unique")) uniq = true;
            if (acceptKw("priority")) prio = true;
            const Token kw = m_lex.next();  // caseStart (verilog.y:3890-3897)
            expectPunct("(");
            AstNodeExpr* const exprp = parseExpr();
            expectPunct(")");
            AstCase* const casep
                = new AstCase{kw.fl,
                              kw.text == "case"    ? VCaseType::CT_CASE
                              : kw.text == "casez" ? VCaseType::CT_CASEZ
                                                   : VCaseType::CT_CASEX,
                              exprp, nullptr};
            while (!isKw("endcase")) {  // case_itemList (verilog.y:3905-3912)
                if (isKw("default")) {
                    const Token dkw = m_lex.next();
                    acceptPunct(":");
                    casep->addItemsp(new AstCaseItem{dkw.fl, nullptr, parseStmt()});
                } else {
                    AstNodeExpr* condsp = parseExpr();
                    while (acceptPunct(",")) condsp->addNext(parseExpr());
                    const Token colon = expectPunct(":");
                    casep->addItemsp(new AstCaseItem{colon.fl, condsp, parseStmt()});
                }
            }
            expectKw("endcase");
            if (uniq) casep->uniquePragma(true);
            if (prio) casep->priorityPragma(true);
            return casep;
        }
        if (t.text == "do") {  // yDO stmt yWHILE '(' expr ')' ';' (verilog.y:3701-3704)
            const Token kw = m_lex.next();
            AstLoop* const loopp = new AstLoop{kw.fl, parseStmt()};
            expectKw("while");
            const Token lp = expectPunct("(");
            loopp->addContsp(new AstLoopTest{lp.fl, loopp, parseExpr()});
            expectPunct(")");
            acceptPunct(";");
            return loopp;
        }
        if (t.text == "while") {  // statement_item: yWHILE '(' expr ')' stmt (verilog.y:3716-3720)
            const Token kw = m_lex.next();
            const Token lp = expectPunct("(");
            AstNodeExpr* const condp = parseExpr();
            expectPunct(")");
            AstLoop* const loopp = new AstLoop{kw.fl};
            loopp->addStmtsp(new AstLoopTest{lp.fl, loopp, condp});
            loopp->addStmtsp(parseStmt());
            return loopp;
        }
        if (t.text == "for") {  // statement_item: yFOR (verilog.y:3698-3706)
            const Token kw = m_lex.next();
            expectPunct("(");
            AstNode* initp = nullptr;
            if (isKw("int") || isKw("integer")) {  // for_initialization with declaration
                varReset(VVarType::VAR, 0);
                setDType(parseDataTypeE());
                m_varDeclTyped = true;
                const Token name = expectId();
                std::vector<AstRange*> dims;
                AstVar* const varp = createVariable(name.fl, name.text, dims);
                varp->lifetime(VLifetime::AUTOMATIC_IMPLICIT);
                const Token eq = expectPunct("=");
                initp = varp;
                initp->addNext(new AstAssign{eq.fl, new AstParseRef{name.fl, name.text}, parseExpr()});
                expectPunct(";");
            } else {
                initp = parseAssignStmt();  // consumes ';'
            }
            AstNodeExpr* const condp = isPunct(";") ? nullptr : parseExpr();
            FileLine* const condFl = condp ? condp->fileline() : kw.fl;
            expectPunct(";");
            AstNode* stepp = nullptr;
            if (!isPunct(")")) stepp = parseAssignNoSemi();
            expectPunct(")");
            AstNodeStmt* const bodyp = parseStmt();
            AstBegin* const blockp = new AstBegin{kw.fl, "", initp, true};
            AstLoop* const loopp = new AstLoop{kw.fl};
            if (condp) loopp->addStmtsp(new AstLoopTest{condFl, loopp, condp});
            loopp->addStmtsp(bodyp);
            loopp->addContsp(stepp);
            blockp->addStmtsp(loopp);
            return blockp;
        }
        if (t.text == "assert" || t.text == "cover") {
            if (isKw("property", 1)) return parseConcurrentAssertion();
            // simple_immediate_assertion_statement (verilog.y:6443-6451)
            const Token kw = m_lex.next();
            expectPunct("(");
            AstNodeExpr* const exprp = parseExpr();
            expectPunct(")");
            AstNode* passp = nullptr;
            AstNode* failp = nullptr;
            if (isKw("else")) {
                m_lex.next();
                failp = parseStmt();
            } else {
                passp = parseStmt();
                if (acceptKw("else")) failp = parseStmt();
            }
            if (kw.text == "cover") return new AstCover{kw.fl, exprp, passp, VAssertType::SIMPLE_IMMEDIATE};
            return new AstAssert{kw.fl, exprp, passp, failp, VAssertType::SIMPLE_IMMEDIATE,
                                 VAssertDirectiveType::ASSERT};
        }
        if (!isReserved(t.text) && isPunct(":", 1)) {  // id ':' statement_item (verilog.y:3575)
            const Token label = m_lex.next();
            m_lex.next();
            return new AstBegin{label.fl, label.text, parseStmt(), false};
        }
        if (!isReserved(t.text)) {
            // task enable: id ';' or id '(' list_of_argumentsE ')' ';' (verilog.y:4133-4134)
            const Token& t1 = m_lex.peek(1);
            if (t1.kind == Tok::PUNCT && (t1.text == ";" || t1.text == "(")) {
                const Token name = m_lex.next();
                AstArg* argsp = nullptr;
                if (acceptPunct("(")) argsp = parseCallArgs();
                expectPunct(";");
                return (new AstTaskRef{name.fl, name.text, argsp})->makeStmt();
            }
            return parseAssignStmt();
        }
        syntaxError(t, "statement");
    }

    // foperator_assignment / fexprLvalue yP_LTE expr (verilog.y:3586-3590, 3796-3797)
    AstNodeStmt* parseAssignNoSemi() {
        AstNodeExpr* const lhsp = parseLvalue();
        if (isPunct("=")) {
            const Token eq = m_lex.next();
            return new AstAssign{eq.fl, lhsp, parseExpr()};
        }
        if (isPunct("<=")) {
            const Token le = m_lex.next();
            return new AstAssignDly{le.fl, lhsp, parseExpr()};
        }
        {  // foperator_assignment (verilog.y:3798-3820)
            static const std::pair<const char*, decltype(AstAssignCompound::Add)> ops[]
                = {{"+=", AstAssignCompound::Add},      {"-=", AstAssignCompound::Sub},
                   {"*=", AstAssignCompound::Mul},      {"/=", AstAssignCompound::Div},
                   {"%=", AstAssignCompound::ModDiv},   {"&=", AstAssignCompound::And},
                   {"|=", AstAssignCompound::Or},       {"^=", AstAssignCompound::Xor},
                   {"<<=", AstAssignCompound::ShiftL},  {">>=", AstAssignCompound::ShiftR},
                   {">>>=", AstAssignCompound::ShiftRS}};
            for (const auto& o : ops) {
                if (isPunct(o.first)) {
                    const Token op = m_lex.next();
                    return new AstAssignCompound{o.second, op.fl, lhsp, parseExpr()};
                }
            }
        }
        if (isPunct("++") || isPunct("--")) {  // finc_or_dec_expression (verilog.y:3594, 3822-3826)
            const Token op = m_lex.next();
            AstNodeExpr* const exprp = op.text == "++" ? static_cast<AstNodeExpr*>(new AstPostInc{op.fl, lhsp})
                                                       : static_cast<AstNodeExpr*>(new AstPostDec{op.fl, lhsp});
            return new AstStmtExpr{lhsp->fileline(), exprp};
        }
        syntaxError(m_lex.peek(), "'=' or '<='");
    }
    AstNodeStmt* parseAssignStmt() {
        AstNodeStmt* const stmtp = parseAssignNoSemi();
        expectPunct(";");
        return stmtp;
    }

    // system_t_stmt_call subset (verilog.y:4213-4330)
    AstNodeStmt* parseSystemTask() {
        const Token name = m_lex.next();
        AstNodeExpr* argsp = nullptr;
        bool hadParen = false;
        if (acceptPunct("(")) {
            hadParen = true;
            if (!isPunct(")")) {
                argsp = parseExpr();
                while (acceptPunct(",")) argsp->addNext(parseExpr());
            }
            expectPunct(")");
        }
        (void)hadParen;
        if (name.text == "$display") return new AstDisplay{name.fl, VDisplayType::DT_DISPLAY, nullptr, argsp};
        if (name.text == "$write") return new AstDisplay{name.fl, VDisplayType::DT_WRITE, nullptr, argsp};
        if (name.text == "$finish") {
            if (argsp) VL_DO_DANGLING(argsp->deleteTree(), argsp);
            return new AstFinish{name.fl};
        }
        if (name.text == "$stop") {
            if (argsp) VL_DO_DANGLING(argsp->deleteTree(), argsp);
            return new AstStop{name.fl, false};
        }
        if (name.text == "$error") {  // yD_ERROR (verilog.y:4319-4322)
            AstNode* const dispp = new AstDisplay{name.fl, VDisplayType::DT_ERROR, nullptr, argsp};
            dispp->addNext(new AstStop{name.fl, false});
            return VN_AS(dispp, NodeStmt);
        }
        if (name.text == "$fatal") {  // yD_FATAL: the first argument is the finish number (verilog.y:4323-4331)
            AstNodeExpr* restp = nullptr;
            if (argsp) {
                restp = VN_CAST(argsp->nextp(), NodeExpr);
                if (restp) restp->unlinkFrBackWithNext();
                VL_DO_DANGLING(argsp->deleteTree(), argsp);
            }
            AstNode* const dispp = new AstDisplay{name.fl, VDisplayType::DT_FATAL, nullptr, restp};
            dispp->addNext(new AstStop{name.fl, true});
            return VN_AS(dispp, NodeStmt);
        }
        if (name.text == "$dumpfile") return new AstDumpCtl{name.fl, VDumpCtlType::FILE, argsp};
        if (name.text == "$dumpvars") {
            return new AstDumpCtl{name.fl, VDumpCtlType::VARS, argsp ? argsp : new AstConst{name.fl, 0}};
        }
        syntaxError(name, "a supported system task");
    }

    // -- expressions --------------------------------------------------------------
    // variable_lvalue / fexprLvalue: exprScope or concatenation (verilog.y:5288-5291)
    AstNodeExpr* parseLvalue() {
        if (isPunct("{")) {
            const Token lb = m_lex.next();
            AstNodeExpr* const catep = parseCateList();
            expectPunct("}");
            return new AstReplicate{lb.fl, catep, 1};
        }
        return parseIdArrayed();
    }
    // idArrayed (verilog.y:6247-6256)
    // list_of_argumentsE: AstArg per positional argument (verilog.y:5462-5476)
    AstArg* parseCallArgs() {
        AstArg* listp = nullptr;
        if (!isPunct(")")) {
            do {
                AstNodeExpr* const exprp = parseExpr();
                AstArg* const argp = new AstArg{exprp->fileline(), "", exprp};
                if (listp) {
                    listp->addNext(argp);
                } else {
                    listp = argp;
                }
            } while (acceptPunct(","));
        }
        expectPunct(")");
        return listp;
    }
    AstNodeExpr* parseIdArrayed() {
        const Token name = expectId();
        if (isPunct("(")) {  // function call: id '(' list_of_argumentsE ')' (verilog.y:4149-4150)
            m_lex.next();
            return parseSelects(new AstFuncRef{name.fl, name.text, parseCallArgs()});
        }
        AstNodeExpr* exprp = new AstParseRef{name.fl, name.text, nullptr, nullptr};
        return parseSelects(exprp);
    }
    AstNodeExpr* parseSelects(AstNodeExpr* exprp) {
        while (isPunct("[")) {
            const Token lb = m_lex.next();
            AstNodeExpr* const firstp = parseExpr();
            if (acceptPunct(":")) {
                exprp = new AstSelExtract{lb.fl, exprp, firstp, parseExpr()};
            } else if (acceptPunct("+:")) {
                exprp = new AstSelPlus{lb.fl, exprp, firstp, parseExpr()};
            } else if (acceptPunct("-:")) {
                exprp = new AstSelMinus{lb.fl, exprp, firstp, parseExpr()};
            } else {
                exprp = new AstSelBit{lb.fl, exprp, firstp};
            }
            expectPunct("]");
        }
        return exprp;
    }
    // cateList (verilog.y:5398-5402)
    AstNodeExpr* parseCateList() {
        AstNodeExpr* exprp = parseExpr();
        while (isPunct(",")) {
            const Token comma = m_lex.next();
            exprp = new AstConcat{comma.fl, exprp, parseExpr()};
        }
        return exprp;
    }

    // Binary operator table, in IEEE 1800-2017 Table 11-2 precedence (the %left/%right lines of
    // verilog.y:1040-1100 encode the same order); node built as the expr rules (:5044-5090) do
    static int binaryPrec(const std::string& op) {
        static const std::map<std::string, int> prec
            = {{"**", 12}, {"*", 11},  {"/", 11},  {"%", 11},   {"+", 10},  {"-", 10},  {"<<", 9},  {">>", 9},
               {"<<<", 9}, {">>>", 9}, {"<", 8},   {"<=", 8},   {">", 8},   {">=", 8},  {"==", 7},  {"!=", 7},
               {"===", 7}, {"!==", 7}, {"==?", 7}, {"!=?", 7},  {"&", 6},   {"^", 5},   {"~^", 5},  {"^~", 5},
               {"|", 4},   {"&&", 3},  {"||", 2}};
        const auto it = prec.find(op);
        return it == prec.end() ? -1 : it->second;
    }
    static AstNodeExpr* newBinary(const Token& op, AstNodeExpr* l, AstNodeExpr* r) {
        FileLine* const fl = op.fl;
        const std::string& s = op.text;
        if (s == "+") return new AstAdd{fl, l, r};
        if (s == "-") return new AstSub{fl, l, r};
        if (s == "*") return new AstMul{fl, l, r};
        if (s == "/") return new AstDiv{fl, l, r};
        if (s == "%") return new AstModDiv{fl, l, r};
        if (s == "**") return new AstPow{fl, l, r};
        if (s == "==") return new AstEq{fl, l, r};
        if (s == "!=") return new AstNeq{fl, l, r};
        if (s == "===") return new AstEqCase{fl, l, r};
        if (s == "!==") return new AstNeqCase{fl, l, r};
        if (s == "==?") return new AstEqWild{fl, l, r};
        if (s == "!=?") return new AstNeqWild{fl, l, r};
        if (s == "&&") return new AstLogAnd{fl, l, r};
        if (s == "||") return new AstLogOr{fl, l, r};
        if (s == "<") return new AstLt{fl, l, r};
        if (s == ">") return new AstGt{fl, l, r};
        if (s == ">=") return new AstGte{fl, l, r};
        if (s == "<=") return new AstLte{fl, l, r};
        if (s == "&") return new AstAnd{fl, l, r};
        if (s == "|") return new AstOr{fl, l, r};
        if (s == "^") return new AstXor{fl, l, r};
        if (s == "~^" || s == "^~") return new AstNot{fl, new AstXor{fl, l, r}};
        if (s == "<<" || s == "<<<") return new AstShiftL{fl, l, r};
        if (s == ">>") return new AstShiftR{fl, l, r};
        if (s == ">>>") return new AstShiftRS{fl, l, r};
        fl->v3fatalSrc("operator table out of sync: " << s);
        return nullptr;
    }
    AstNodeExpr* parseBinary(int minPrec) {
        AstNodeExpr* lhsp = parseUnary();
        for (;;) {
            const Token& t = m_lex.peek();
            if (t.kind != Tok::PUNCT) break;
            const int prec = binaryPrec(t.text);
            if (prec < 0 || prec < minPrec) break;
            const Token op = m_lex.next();
            // all left associative except '**' which the grammar also declares %left
            AstNodeExpr* const rhsp = parseBinary(prec + 1);
            lhsp = newBinary(op, lhsp, rhsp);
        }
        return lhsp;
    }
    // expr '?' expr ':' expr, right associative (verilog.y:5087)
    AstNodeExpr* parseExpr() {
        AstNodeExpr* const condp = parseBinary(0);
        if (isPunct("?")) {
            const Token q = m_lex.next();
            AstNodeExpr* const thenp = parseExpr();
            expectPunct(":");
            AstNodeExpr* const elsep = parseExpr();
            return new AstCond{q.fl, condp, thenp, elsep};
        }
        return condp;
    }
    // unary operators (verilog.y:4994-5003)
    AstNodeExpr* parseUnary() {
        const Token& t = m_lex.peek();
        if (t.kind == Tok::PUNCT) {
            const std::string s = t.text;
            if (s == "+") {
                m_lex.next();
                return parseUnary();
            }
            if (s == "-" || s == "!" || s == "~" || s == "&" || s == "|" || s == "^" || s == "~&" || s == "~|"
                || s == "~^" || s == "^~") {
                const Token op = m_lex.next();
                AstNodeExpr* const p = parseUnary();
                if (s == "-") return new AstNegate{op.fl, p};
                if (s == "!") return new AstLogNot{op.fl, p};
                if (s == "~") return new AstNot{op.fl, p};
                if (s == "&") return new AstRedAnd{op.fl, p};
                if (s == "|") return new AstRedOr{op.fl, p};
                if (s == "^") return new AstRedXor{op.fl, p};
                if (s == "~&") return new AstLogNot{op.fl, new AstRedAnd{op.fl, p}};
                if (s == "~|") return new AstLogNot{op.fl, new AstRedOr{op.fl, p}};
                return new AstLogNot{op.fl, new AstRedXor{op.fl, p}};
            }
        }
        return parsePrimary();
    }
    AstNodeExpr* parsePrimary() {
        const Token& t = m_lex.peek();
        switch (t.kind) {
        case Tok::NUM: {  // yaINTNUM (verilog.y:5118): the literal text goes to V3Number
            const Token num = m_lex.next();
            const V3Number value{num.fl, V3Number::VerilogNumberLiteral{}, num.text.c_str()};
            return new AstConst{num.fl, value};
        }
        case Tok::STR: {  // strAsInt (verilog.y:6281-6285)
            const Token str = m_lex.next();
            std::string errMsg;
            const std::string text
                = VString::unquoteSVString(str.text.substr(1, str.text.size() - 2), errMsg);
            if (!errMsg.empty()) str.fl->v3error(errMsg);
            return new AstConst{str.fl, AstConst::VerilogStringLiteral{}, text};
        }
        case Tok::SYSID: return parseSystemFunc();
        case Tok::ID:
            if (isReserved(t.text)) syntaxError(t, "expression");
            return parseIdArrayed();
        case Tok::PUNCT:
            if (t.text == "(") {
                m_lex.next();
                AstNodeExpr* const exprp = parseExpr();
                expectPunct(")");
                return exprp;
            }
            if (t.text == "{") {
                const Token lb = m_lex.next();
                AstNodeExpr* const firstp = parseExpr();
                if (isPunct("{")) {  // '{' constExpr '{' cateList '}' '}' (verilog.y:5123-5124)
                    const Token lb2 = m_lex.next();
                    AstNodeExpr* const catep = parseCateList();
                    expectPunct("}");
                    expectPunct("}");
                    return parseSelects(new AstReplicate{lb2.fl, catep, firstp});
                }
                AstNodeExpr* exprp = firstp;
                while (isPunct(",")) {
                    const Token comma = m_lex.next();
                    exprp = new AstConcat{comma.fl, exprp, parseExpr()};
                }
                expectPunct("}");
                return parseSelects(new AstReplicate{lb.fl, exprp, 1});
            }
            syntaxError(t, "expression");
        default: syntaxError(t, "expression");
        }
    }
    // system_f_call subset (verilog.y:4426-4560)
    AstNodeExpr* parseSystemFunc() {
        const Token name = m_lex.next();
        std::vector<AstNodeExpr*> args;
        if (acceptPunct("(")) {
            if (!isPunct(")")) {
                do { args.push_back(parseExpr()); } while (acceptPunct(","));
            }
            expectPunct(")");
        }
        const auto need = [&](size_t n) {
            if (args.size() != n) syntaxError(name, std::to_string(n) + " argument(s)");
        };
        if (name.text == "$time") {
            need(0);
            return new AstTime{name.fl, VTimescale{VTimescale::NONE}};
        }
        if (name.text == "$urandom" || name.text == "$random") {  // yD_URANDOM / yD_RANDOM parenE (verilog.y:4516-4559)
            need(0);
            return new AstRand{name.fl, nullptr, name.text == "$urandom"};
        }
        if (name.text == "$signed") { need(1); return new AstSigned{name.fl, args[0]}; }
        if (name.text == "$unsigned") { need(1); return new AstUnsigned{name.fl, args[0]}; }
        if (name.text == "$clog2") { need(1); return new AstCLog2{name.fl, args[0]}; }
        if (name.text == "$countones") { need(1); return new AstCountOnes{name.fl, args[0]}; }
        if (name.text == "$onehot") { need(1); return new AstOneHot{name.fl, args[0]}; }
        if (name.text == "$onehot0") { need(1); return new AstOneHot0{name.fl, args[0]}; }
        if (name.text == "$test$plusargs") { need(1); return new AstTestPlusArgs{name.fl, args[0]}; }
        if (name.text == "$bits") { need(1); return new AstAttrOf{name.fl, VAttrType::DIM_BITS, args[0]}; }
        syntaxError(name, "a supported system function");
    }

public:
    Parser(Lexer& lex, AstNetlist* rootp, bool inLibrary, const std::string& libname)
        : m_lex{lex}
        , m_rootp{rootp}
        , m_inLibrary{inLibrary}
        , m_libname{libname} {}
    void parseDescriptions() {
        try {
            while (m_lex.peek().kind != Tok::END) {
                if (isKw("module")) {
                    parseModule();
                } else {
                    syntaxError(m_lex.peek(), "'module'");
                }
            }
        } catch (const SyntaxError&) {}  // reported through v3error; V3Error::abortIfErrors follows
    }
};

// The object behind V3Parse::m_impp in this build
struct ParseLiteImp final {
    AstNetlist* rootp;
    VInFilter* filterp;
    std::string text;  // preprocessed text pushed by V3PreShell::preproc
};

}  // namespace

//======================================================================
// The symbols the generated parser normally provides

V3Parse::V3Parse(AstNetlist* rootp, VInFilter* filterp) {
    m_impp = reinterpret_cast<V3ParseImp*>(new ParseLiteImp{rootp, filterp, ""});
}
V3Parse::~V3Parse() { VL_DO_CLEAR(delete reinterpret_cast<ParseLiteImp*>(m_impp), m_impp = nullptr); }

void V3Parse::parseFile(FileLine* fileline, const string& modname, bool inLibrary, bool inLibMap,
                        const string& libname, const string& errmsg, const std::string& notFoundName) {
    ParseLiteImp* const impp = reinterpret_cast<ParseLiteImp*>(m_impp);
    if (inLibMap) return;
    // Control (.vlt) files and the std package need the full grammar; the oracle build runs
    // with --no-std --no-std-waiver so none are requested
    const string ext = V3Os::filenameNonDir(modname);
    if (ext.size() > 4 && ext.compare(ext.size() - 4, 4, ".vlt") == 0) return;
    impp->text.clear();
    const bool ok = V3PreShell::preproc(fileline, modname, impp->filterp, m_impp, errmsg);
    if (!ok) {
        if (errmsg != "") return;  // Threw error already
        // Fake node for later error reporting, as V3ParseImp::parseFile does (V3ParseImp.cpp:393-400)
        AstNodeModule* const nodep = new AstNotFoundModule{
            fileline, notFoundName.empty() ? V3Os::filenameNonDirExt(modname) : notFoundName, libname};
        v3Global.rootp()->addModulesp(nodep);
        return;
    }
    MacroExpander macros;
    Lexer lexer{fileline, macros.run(impp->text, fileline->filename())};
    Parser parser{lexer, impp->rootp, inLibrary, libname};
    parser.parseDescriptions();
}

void V3Parse::ppPushText(V3ParseImp* impp, const string& text) {
    reinterpret_cast<ParseLiteImp*>(impp)->text += text;
}
void V3Parse::candidatePli(VSpellCheck*) {}

V3PreProc* V3PreProc::createPreProc(FileLine*) { return new PreProcLite; }
void V3PreProc::selfTest() {}
