// rc_frontend.cpp — the data formats either side of the likelihood path, host only.
//
//   model template DSL      src/RarecoalLib/ModelTemplate.hs:78-145   (BRANCHES / EVENTS / CONSTRAINTS)
//   getParamNames           ModelTemplate.hs:147-174
//   instantiateModel        ModelTemplate.hs:176-241  (K = Join then SetPopSize at the same time; constraints)
//   parameter files         ModelTemplate.hs:243-281  (maxl two-column / mcmc '#' style, -X overrides, defaults)
//   histogram file          exampleData/fourPop.histogram.txt:1-5 (parser lives in the un-vendored package
//                           `sequence-formats`; format inferred, parity unpinned by the reference)
//   histogram filters       src/RarecoalLib/Utils.hs:130-177 (maxAf, minAf, conditionOn, excludePatterns,
//                           computeStandardOrder) and branch mapping Utils.hs:114-119,199-206
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/rarecoal_b200.h"
#include "rc_plan.h"

namespace {

struct ParamSpec {
    bool fixed = true;
    double value = 0.0;
    std::string name;
};

struct MTEvent {
    char cmd = 'P';              // P, J, K, S, F, G
    ParamSpec t;
    std::string a, b;            // branch / to, from
    ParamSpec v;                 // size / rate
    bool flag = false;           // F
};

struct MTConstraint {
    ParamSpec lhs, rhs;
    bool greater = true;
};

struct Template {
    std::vector<std::string> branches;
    std::vector<MTEvent> events;
    std::vector<MTConstraint> constraints;
};

struct Parser {
    const std::string& s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string& str) : s(str) {}
    void spaces() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool spaces1() {
        size_t j = i;
        spaces();
        if (i == j) return fail("expected white space");
        return true;
    }
    bool fail(const std::string& m) {
        if (err.empty()) err = m + " at offset " + std::to_string(i);
        return false;
    }
    bool lit(const char* w) {
        size_t n = std::strlen(w);
        if (s.compare(i, n, w) != 0) return fail(std::string("expected \"") + w + "\"");
        i += n;
        return true;
    }
    bool ch(char c) {
        if (i >= s.size() || s[i] != c) return fail(std::string("expected '") + c + "'");
        ++i;
        return true;
    }
    bool ident(std::string& out) {       // many1 (alphaNum | '_')
        size_t j = i;
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_')) ++i;
        if (i == j) return fail("expected a name");
        out = s.substr(j, i - j);
        return true;
    }
    bool param(ParamSpec& p) {           // try floating <|> '<' name '>'
        if (i < s.size() && s[i] == '<') {
            ++i;
            p.fixed = false;
            if (!ident(p.name)) return false;
            return ch('>');
        }
        const char* b = s.c_str() + i;
        char* e = nullptr;
        if (i >= s.size() || !(std::isdigit((unsigned char)*b))) return fail("expected a number or <name>");
        double v = std::strtod(b, &e);
        if (e == b) return fail("expected a number");
        p.fixed = true;
        p.value = v;
        i += (size_t)(e - b);
        return true;
    }
    bool header(const char* word) {      // WORD spaces '=' spaces '[' spaces
        if (!lit(word)) return false;
        spaces();
        if (!ch('=')) return false;
        spaces();
        if (!ch('[')) return false;
        spaces();
        return true;
    }
};

bool parse_template(const std::string& text, Template& t, std::string& err) {
    Parser p(text);
    p.spaces();
    if (!p.header("BRANCHES")) { err = p.err; return false; }
    for (;;) {
        std::string n;
        if (!p.ident(n)) { err = p.err; return false; }
        t.branches.push_back(n);
        p.spaces();
        if (p.i < text.size() && text[p.i] == ',') { ++p.i; p.spaces(); continue; }
        break;
    }
    if (!p.ch(']')) { err = p.err; return false; }
    p.spaces();
    if (!p.header("EVENTS")) { err = p.err; return false; }
    while (p.i < text.size() && text[p.i] != ']') {
        MTEvent e;
        e.cmd = text[p.i];
        // G t branch rate: SetGrowthRate, an extension of the reference grammar (include/rarecoal_b200.h, RC_EV_GROWTH)
        if (std::strchr("PJKSFG", e.cmd) == nullptr) { p.fail("unknown event command"); err = p.err; return false; }
        ++p.i;
        bool ok = p.spaces1() && p.param(e.t) && p.spaces1() && p.ident(e.a);
        if (ok && (e.cmd == 'P' || e.cmd == 'G')) ok = p.spaces1() && p.param(e.v);
        if (ok && e.cmd == 'J') ok = p.spaces1() && p.ident(e.b);
        if (ok && (e.cmd == 'K' || e.cmd == 'S')) ok = p.spaces1() && p.ident(e.b) && p.spaces1() && p.param(e.v);
        if (ok && e.cmd == 'F') {
            ok = p.spaces1();
            if (ok && text.compare(p.i, 4, "True") == 0) { e.flag = true; p.i += 4; }
            else if (ok && text.compare(p.i, 5, "False") == 0) { e.flag = false; p.i += 5; }
            else ok = p.fail("expected True or False");
        }
        if (!ok) { err = p.err; return false; }
        t.events.push_back(e);
        p.spaces();
        if (p.i < text.size() && text[p.i] == ';') { ++p.i; p.spaces(); continue; }
        break;
    }
    if (!p.ch(']')) { err = p.err; return false; }
    p.spaces();
    if (p.i < text.size() && text.compare(p.i, 11, "CONSTRAINTS") == 0) {
        if (!p.header("CONSTRAINTS")) { err = p.err; return false; }
        while (p.i < text.size() && text[p.i] != ']') {
            MTConstraint c;
            bool ok = p.ch('C') && p.spaces1() && p.param(c.lhs);
            if (ok) {
                p.spaces();
                if (p.i < text.size() && (text[p.i] == '<' || text[p.i] == '>')) { c.greater = text[p.i] == '>'; ++p.i; }
                else ok = p.fail("expected < or >");
            }
            if (ok) { p.spaces(); ok = p.param(c.rhs); }
            if (!ok) { err = p.err; return false; }
            t.constraints.push_back(c);
            p.spaces();
            if (p.i < text.size() && text[p.i] == ';') { ++p.i; p.spaces(); continue; }
            break;
        }
        if (!p.ch(']')) { err = p.err; return false; }
        p.spaces();
    }
    return true;
}

// getParamNames (ModelTemplate.hs:147-174): template order, time slot before value slot, a repeated name at its
// LAST appearance (reverse . nub over the reversed list).
std::vector<std::string> param_names(const Template& t) {
    std::vector<std::string> seq;
    for (const MTEvent& e : t.events) {
        if (!e.t.fixed) seq.push_back(e.t.name);
        if (e.cmd != 'J' && e.cmd != 'F' && !e.v.fixed) seq.push_back(e.v.name);
    }
    std::vector<std::string> out;
    for (size_t i = 0; i < seq.size(); ++i) {
        bool later = false;
        for (size_t j = i + 1; j < seq.size() && !later; ++j) later = seq[j] == seq[i];
        if (!later) out.push_back(seq[i]);
    }
    return out;
}

std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') o.push_back('\\');
        o.push_back(c);
    }
    return o;
}

std::string spec_json(const ParamSpec& p) {
    char buf[64];
    if (p.fixed) { std::snprintf(buf, sizeof buf, "{\"fixed\": %.17g}", p.value); return buf; }
    return "{\"var\": \"" + json_escape(p.name) + "\"}";
}

void copy_out(const std::string& s, char* out, int cap) {
    if (out && cap > 0) std::snprintf(out, (size_t)cap, "%s", s.c_str());
}

struct Hist {
    std::vector<std::string> names;
    std::vector<int> nVec;
    int minM = 1, maxM = 0;
    long long total = 0;
    std::vector<int> conditionOn;
    std::vector<std::vector<int>> exclude;
    std::vector<std::vector<int>> pats;
    std::vector<long long> counts;
};

std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == sep) { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}

}  // namespace

struct rc_hist {
    Hist h;
};

extern "C" {

int rc_template_to_json(const char* text, char* out, int cap, char* err, int errCap) {
    if (!text) return RC_ERR_ARG;
    Template t;
    std::string e;
    if (!parse_template(text, t, e)) { copy_out(e, err, errCap); return -RC_ERR_TEMPLATE; }
    std::ostringstream o;
    o << "{\"branches\": [";
    for (size_t i = 0; i < t.branches.size(); ++i) o << (i ? ", " : "") << '"' << json_escape(t.branches[i]) << '"';
    o << "], \"events\": [";
    for (size_t i = 0; i < t.events.size(); ++i) {
        const MTEvent& ev = t.events[i];
        o << (i ? ", " : "") << "{\"cmd\": \"" << ev.cmd << "\", \"t\": " << spec_json(ev.t) << ", \"a\": \"" << json_escape(ev.a)
          << "\", \"b\": \"" << json_escape(ev.b) << "\", \"v\": " << spec_json(ev.v) << ", \"flag\": " << (ev.flag ? "true" : "false") << "}";
    }
    o << "], \"constraints\": [";
    for (size_t i = 0; i < t.constraints.size(); ++i)
        o << (i ? ", " : "") << "{\"lhs\": " << spec_json(t.constraints[i].lhs) << ", \"rhs\": " << spec_json(t.constraints[i].rhs)
          << ", \"op\": \"" << (t.constraints[i].greater ? ">" : "<") << "\"}";
    o << "], \"params\": [";
    std::vector<std::string> pn = param_names(t);
    for (size_t i = 0; i < pn.size(); ++i) o << (i ? ", " : "") << '"' << json_escape(pn[i]) << '"';
    o << "]}";
    std::string s = o.str();
    if ((int)s.size() + 1 > cap) return (int)s.size() + 1;   // caller retries with a larger buffer
    copy_out(s, out, cap);
    return 0;
}

// instantiateModel for one parameter vector on a parsed template (ModelTemplate.hs:176-241): 0, or -RC_ERR_MODEL / -RC_ERR_TEMPLATE
// with a message
static int instantiate_one(const Template& t, int nParams, const char* const* names, const double* values, std::vector<rc_event>& evs,
                           std::string& msg) {
    bool missing = false;
    std::string missingName;
    auto get = [&](const ParamSpec& p) -> double {
        if (p.fixed) return p.value;
        for (int i = 0; i < nParams; ++i)            // `lookup`: first match
            if (p.name == names[i]) return values[i];
        if (!missing) { missing = true; missingName = p.name; }
        return 0.0;
    };
    auto branch = [&](const std::string& n) -> int {
        for (size_t i = 0; i < t.branches.size(); ++i) if (t.branches[i] == n) return (int)i;
        return -1;
    };
    evs.clear();
    for (const MTEvent& m : t.events) {
        int a = branch(m.a), b = m.b.empty() ? 0 : branch(m.b);
        if (a < 0 || b < 0) {
            msg = "did not find model-branch name \"" + (a < 0 ? m.a : m.b) + "\"";
            return -RC_ERR_MODEL;                     // getModelBranchIndex, ModelTemplate.hs:215-219
        }
        double tt = get(m.t);
        rc_event ev;
        ev.t = tt;
        ev.k = a;
        ev.l = 0;
        ev.v = 0.0;
        switch (m.cmd) {
            case 'P': ev.type = RC_EV_POPSIZE; ev.v = get(m.v); evs.push_back(ev); break;
            case 'J': ev.type = RC_EV_JOIN; ev.l = b; evs.push_back(ev); break;
            case 'K':                                   // Join, then SetPopSize at the same time (ModelTemplate.hs:197-203)
                ev.type = RC_EV_JOIN; ev.l = b; evs.push_back(ev);
                ev.type = RC_EV_POPSIZE; ev.l = 0; ev.v = get(m.v); evs.push_back(ev);
                break;
            case 'S': ev.type = RC_EV_SPLIT; ev.l = b; ev.v = get(m.v); evs.push_back(ev); break;
            case 'G': ev.type = RC_EV_GROWTH; ev.v = get(m.v); evs.push_back(ev); break;
            default: ev.type = RC_EV_FREEZE; ev.v = m.flag ? 1.0 : 0.0; evs.push_back(ev); break;
        }
    }
    if (missing) {
        msg = "Error in Template: could not find parameter \"" + missingName + "\"";
        return -RC_ERR_TEMPLATE;                      // RarecoalModeltemplateException, ModelTemplate.hs:226
    }
    for (const MTConstraint& c : t.constraints) {     // validateConstraints, ModelTemplate.hs:229-241
        double p1 = get(c.lhs), p2 = get(c.rhs);
        if (missing) { msg = "Error in Template: could not find parameter \"" + missingName + "\""; return -RC_ERR_TEMPLATE; }
        if ((c.greater && p1 <= p2) || (!c.greater && p1 >= p2)) {
            msg = "constraint violated";
            return -RC_ERR_MODEL;
        }
    }
    return 0;
}

int rc_template_instantiate(const char* text, int nParams, const char* const* names, const double* values, rc_event* out,
                            int cap, char* err, int errCap) {
    if (!text || nParams < 0 || (nParams > 0 && (!names || !values))) return RC_ERR_ARG;
    Template t;
    std::string e;
    if (!parse_template(text, t, e)) { copy_out(e, err, errCap); return -RC_ERR_TEMPLATE; }
    std::vector<rc_event> evs;
    const int rc = instantiate_one(t, nParams, names, values, evs, e);
    if (rc < 0) { copy_out(e, err, errCap); return rc; }
    if (out)
        for (size_t i = 0; i < evs.size() && (int)i < cap; ++i) out[i] = evs[i];
    return (int)evs.size();
}

// The same for B parameter vectors (values[B][nParams]) with the template parsed once: the events of the models one after the
// other in `out` (evOff[b] .. evOff[b + 1]), per model a status (0, RC_ERR_MODEL or RC_ERR_TEMPLATE: no events) and, if
// `penalty` is given, the regularisation penalty (rc_reg_penalty).  What an optimiser or sampler needs per step as one call;
// returns the number of events or a negative code (template does not parse, `out` too small).
int rc_template_instantiate_batch(const char* text, int nParams, const char* const* names, int B, const double* values, double reg,
                                  rc_event* out, int cap, int32_t* evOff, int32_t* status, double* penalty, char* err, int errCap) {
    if (!text || nParams < 0 || B < 0 || (nParams > 0 && (!names || !values)) || !out || !evOff || !status) return RC_ERR_ARG;
    Template t;
    std::string e;
    if (!parse_template(text, t, e)) { copy_out(e, err, errCap); return -RC_ERR_TEMPLATE; }
    std::vector<rc_event> evs;
    int n = 0;
    evOff[0] = 0;
    for (int b = 0; b < B; ++b) {
        std::string msg;
        const int rc = instantiate_one(t, nParams, names, values + (size_t)b * nParams, evs, msg);
        status[b] = rc < 0 ? -rc : 0;
        if (rc < 0) evs.clear();
        if (n + (int)evs.size() > cap) { copy_out("rc_template_instantiate_batch: event buffer too small", err, errCap); return RC_ERR_ARG; }
        for (const rc_event& ev : evs) out[n++] = ev;
        evOff[b + 1] = n;
        if (penalty) penalty[b] = rc < 0 ? 0.0 : rc_reg_penalty((int)t.branches.size(), reg, out + evOff[b], (int)evs.size());
        if (rc < 0 && e.empty()) e = msg;
    }
    if (!e.empty()) copy_out(e, err, errCap);
    return n;
}

// makeParameterDict (ModelTemplate.hs:243-263): text of a maxl (name value) or mcmc ('#' first, skip 3 lines, MaxL
// column) output file -> "name value" pairs in file order.  Returns the number of pairs.
int rc_params_parse(const char* text, char* namesOut, int namesCap, double* values, int cap) {
    if (!text) return RC_ERR_ARG;
    std::vector<std::string> lines = split(text, '\n');
    std::vector<std::pair<std::string, double>> kv;
    const bool mcmc = !lines.empty() && !lines[0].empty() && lines[0][0] == '#';
    for (size_t li = mcmc ? 3 : 0; li < lines.size(); ++li) {
        std::istringstream ls(lines[li]);
        std::vector<std::string> w;
        std::string tok;
        while (ls >> tok) w.push_back(tok);
        if (mcmc ? w.size() >= 2 : w.size() == 2) {
            char* e = nullptr;
            double v = std::strtod(w[1].c_str(), &e);
            if (e != w[1].c_str()) kv.emplace_back(w[0], v);
        }
    }
    std::string names;
    for (size_t i = 0; i < kv.size(); ++i) {
        names += (i ? "\n" : "") + kv[i].first;
        if (values && (int)i < cap) values[i] = kv[i].second;
    }
    copy_out(names, namesOut, namesCap);
    return (int)kv.size();
}

// ---- histogram ---------------------------------------------------------------------------------------------------
rc_hist* rc_hist_read(const char* path, char* err, int errCap) {
    std::ifstream f(path ? path : "");
    if (!f) { copy_out(std::string("cannot open histogram ") + (path ? path : ""), err, errCap); return nullptr; }
    rc_hist* H = new rc_hist();
    Hist& h = H->h;
    std::string line;
    bool haveMax = false, haveN = false;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        size_t eq = line.find('=');
        if (eq != std::string::npos && std::isalpha((unsigned char)line[0])) {
            std::string key = line.substr(0, eq), val = line.substr(eq + 1);
            if (key == "NAMES") h.names = split(val, ',');
            else if (key == "N") { for (auto& s : split(val, ',')) h.nVec.push_back(std::atoi(s.c_str())); haveN = true; }
            else if (key == "MAX_M") { h.maxM = std::atoi(val.c_str()); haveMax = true; }
            else if (key == "MIN_M") h.minM = std::atoi(val.c_str());
            else if (key == "TOTAL_SITES") h.total = std::atoll(val.c_str());
            else if (key == "CONDITION_ON") { for (auto& s : split(val, ',')) if (!s.empty()) h.conditionOn.push_back(std::atoi(s.c_str())); }
            else if (key == "EXCLUDE_PATTERNS") {
                for (auto& ps : split(val, ';')) {
                    if (ps.empty()) continue;
                    std::vector<int> p;
                    for (auto& s : split(ps, ',')) p.push_back(std::atoi(s.c_str()));
                    h.exclude.push_back(p);
                }
            }
            continue;
        }
        std::istringstream ls(line);
        std::string pat;
        long long c = 0;
        if (!(ls >> pat >> c)) { copy_out("malformed histogram row: " + line, err, errCap); delete H; return nullptr; }
        std::vector<int> p;
        for (auto& s : split(pat, ',')) p.push_back(std::atoi(s.c_str()));
        h.pats.push_back(p);
        h.counts.push_back(c);
    }
    if (!haveN || !haveMax) { copy_out("histogram header needs N= and MAX_M=", err, errCap); delete H; return nullptr; }
    if (h.names.empty()) for (size_t i = 0; i < h.nVec.size(); ++i) h.names.push_back("pop" + std::to_string(i));
    for (auto& p : h.pats)
        if (p.size() != h.nVec.size()) { copy_out("histogram row with a wrong number of populations", err, errCap); delete H; return nullptr; }
    return H;
}

void rc_hist_free(rc_hist* h) { delete h; }

int rc_hist_info(const rc_hist* H, int32_t* nPops, int32_t* nRows, int32_t* minM, int32_t* maxM, int64_t* totalSites,
                 int32_t* nVec, char* names, int namesCap) {
    if (!H) return RC_ERR_ARG;
    const Hist& h = H->h;
    if (nPops) *nPops = (int)h.nVec.size();
    if (nRows) *nRows = (int)h.pats.size();
    if (minM) *minM = h.minM;
    if (maxM) *maxM = h.maxM;
    if (totalSites) *totalSites = h.total;
    if (nVec) for (size_t i = 0; i < h.nVec.size(); ++i) nVec[i] = h.nVec[i];
    std::string n;
    for (size_t i = 0; i < h.names.size(); ++i) n += (i ? "," : "") + h.names[i];
    copy_out(n, names, namesCap);
    return RC_OK;
}

// loadHistogram's filters (Utils.hs:130-167,184-185) + computeStandardOrder (:169-177) + mapping of patterns and
// nVec to model-branch order (:114-119; a model branch absent from the histogram is a ghost: n = 0, m = 0).
// maxAf == 0 keeps the file's MAX_M.  Returns the number of patterns (standard order) and fills the outputs up to cap
// rows; negative on error (RC_ERR_HIST).
int rc_hist_problem(const rc_hist* H, int maxAf, int minAf, const int32_t* conditionOn, int nCond, const int32_t* exclude,
                    int nExclude, int nModelBranches, const char* const* modelBranches, int32_t* patternsOut,
                    int64_t* countsOut, int32_t* nVecOut, int cap, char* err, int errCap) {
    if (!H || nModelBranches < 1 || !modelBranches) return RC_ERR_ARG;
    const Hist& h = H->h;
    const int P = (int)h.nVec.size();
    int useMax = maxAf == 0 ? h.maxM : maxAf;
    if (useMax > h.maxM || useMax < h.minM) { copy_out("illegal maxAF", err, errCap); return -RC_ERR_HIST; }      // Utils.hs:153
    if (minAf > useMax || minAf < h.minM) { copy_out("illegal minAf", err, errCap); return -RC_ERR_HIST; }         // Utils.hs:162
    // filterConditionOn / filterExcludePatterns (Utils.hs:129-149) REPLACE raConditionOn / raExcludePatterns when the option
    // is given; the lists of the file's header are kept only when the option is empty
    std::vector<int> cond;
    if (nCond > 0) {
        for (int i = 0; i < nCond; ++i) {
            if (conditionOn[i] < 0 || conditionOn[i] >= P) { copy_out("illegal conditionOn indices", err, errCap); return -RC_ERR_HIST; }
            cond.push_back(conditionOn[i]);
        }
    } else cond = h.conditionOn;
    std::vector<std::vector<int>> excl;
    if (nExclude > 0)
        for (int i = 0; i < nExclude; ++i) excl.emplace_back(exclude + (size_t)i * P, exclude + (size_t)(i + 1) * P);
    else excl = h.exclude;
    // validateBranchNameCongruency (Utils.hs:199-206)
    std::vector<int> src((size_t)nModelBranches, -1);
    for (int b = 0; b < nModelBranches; ++b)
        for (int q = 0; q < P; ++q) if (h.names[(size_t)q] == modelBranches[b]) { src[(size_t)b] = q; break; }   // elemIndex: first match (Utils.hs:118)
    for (int q = 0; q < P; ++q) {
        bool found = false;
        for (int b = 0; b < nModelBranches; ++b) found = found || h.names[(size_t)q] == modelBranches[b];
        if (!found) { copy_out("histogram branch \"" + h.names[(size_t)q] + "\" not found in model branches", err, errCap); return -RC_ERR_HIST; }
    }
    std::map<std::vector<int>, long long> counts;
    for (size_t i = 0; i < h.pats.size(); ++i) counts[h.pats[i]] += h.counts[i];
    std::vector<std::vector<int>> order = rc::all_configs(useMax, h.nVec);
    int n = 0;
    for (const auto& p : order) {
        int sum = 0;
        for (int v : p) sum += v;
        bool keep = sum >= minAf;
        for (int c : cond) keep = keep && p[(size_t)c] > 0;
        for (const auto& x : excl) keep = keep && p != x;
        if (!keep) continue;
        if (n < cap) {
            if (patternsOut)
                for (int b = 0; b < nModelBranches; ++b) patternsOut[(size_t)n * nModelBranches + b] = src[(size_t)b] < 0 ? 0 : p[(size_t)src[(size_t)b]];
            if (countsOut) {
                auto it = counts.find(p);
                countsOut[n] = it == counts.end() ? 0 : it->second;      // Map.findWithDefault 0 (MaxUtils.hs:112)
            }
        }
        ++n;
    }
    if (nVecOut)
        for (int b = 0; b < nModelBranches; ++b) nVecOut[b] = src[(size_t)b] < 0 ? 0 : h.nVec[(size_t)src[(size_t)b]];
    return n;
}

}  // extern "C"
