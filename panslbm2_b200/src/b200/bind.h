// Glue between the drop-in PANSLBM2 headers (src/particle, src/equation, src/utility of this tree) and the C-ABI of
// libpanslbm_b200.so (include/panslbm_c.h).  Nothing numerical happens on the host: lambdas are evaluated on the boundary
// planes once and baked into device arrays (pl_bc), everything else is forwarded to the host-pointer surface (plh_*),
// which mirrors the caller's arrays on the device and fuses the calls of a time loop into one pass per step.
//
// Build line of a reference program against this tree (the reference's own is README.md:23-25):
//     g++ -O2 production/heatsink3D.cpp -I<repo>/include -L<repo>/panslbm2_b200 -lpanslbm_b200 -Wl,-rpath,<repo>/panslbm2_b200
#pragma once
#if __has_include("../../../include/panslbm_c.h")
#include "../../../include/panslbm_c.h"
#else
#include <panslbm_c.h>
#endif

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>
#include <unistd.h>

// ---- allocation hook -------------------------------------------------------------------------------------------------
// The drivers allocate their fields with `new double[nxyz]` (production/heatsink3D.cpp:50-59) and std::vector, hand the raw
// pointers to every call and read them directly in between.  Large blocks therefore come from plh_alloc: ordinary host
// memory the runtime can mirror on the device and keep coherent lazily.  Define PANSLBM_B200_NO_ALLOC_HOOK in all but one
// translation unit of a multi-file program (replacement functions must be defined once), or everywhere to opt out
// (arrays are then staged through the device around every call: correct, slow).
#ifndef PANSLBM_B200_ALLOC_THRESHOLD
#define PANSLBM_B200_ALLOC_THRESHOLD 4096
#endif
#ifndef PANSLBM_B200_NO_ALLOC_HOOK
namespace PANSLBM2 { namespace b200 {
    inline void* hooked_new(std::size_t n) {
        void* p = n >= PANSLBM_B200_ALLOC_THRESHOLD ? plh_alloc(n) : std::malloc(n ? n : 1);
        if (!p) throw std::bad_alloc();
        return p;
    }
    inline void hooked_delete(void* p) noexcept {
        if (!p) return;
        if (plh_owns(p)) plh_free(p); else std::free(p);
    }
} }
void* operator new(std::size_t n) { return PANSLBM2::b200::hooked_new(n); }
void* operator new[](std::size_t n) { return PANSLBM2::b200::hooked_new(n); }
void operator delete(void* p) noexcept { PANSLBM2::b200::hooked_delete(p); }
void operator delete[](void* p) noexcept { PANSLBM2::b200::hooked_delete(p); }
void operator delete(void* p, std::size_t) noexcept { PANSLBM2::b200::hooked_delete(p); }
void operator delete[](void* p, std::size_t) noexcept { PANSLBM2::b200::hooked_delete(p); }
#endif

namespace PANSLBM2 {
    namespace {     // bounce-back plane types (d3q15.h:19-22, d2q9.h:19-22); defined once here so that both lattices fit in one program
        const int BARRIER = 1;
        const int MIRROR = 2;
    }
namespace b200 {
    // The reference reports nothing (assert only, d3q15.h:37); a failing device call cannot be ignored, so it is fatal and loud.
    inline void check(int rc, const char* where) {
        if (rc == 0) return;
        const char* a = plh_last_error();
        std::fprintf(stderr, "panslbm_b200: %s failed: %s%s%s\n", where, a ? a : "", (a && *a) ? " | " : "", pl_last_error());
        std::abort();
    }

    struct none_t {};     // absent value callable

    // State every lattice object carries besides the reference's public members.
    struct Core {
        pl_lattice* h = nullptr;
        unsigned long long gen = 0;                 // unique per constructed lattice: keys the baked-plane caches
        std::vector<pl_bc*> planes;                 // baked planes owned by this lattice
        std::vector<std::string> contents;          // their content keys (dedupe: equal planes share one pl_bc)
        std::vector<char> exclusive;                // plane of one volatile call site (updated in place): never shared by content
        std::vector<pl_filter*> filters;            // baked filter weight tables owned by this lattice
        static unsigned long long next_gen() { static unsigned long long g = 0; return ++g; }
        static std::vector<unsigned long long>& live() { static std::vector<unsigned long long> v; return v; }
        void create(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz, double** f0, double** f) {
#ifndef _USE_AVX_DEFINES
            // this program is built like production/nsopt.cpp:2 — the reference would run its scalar templates at every site
            check(pl_set_scalar_order(1), "pl_set_scalar_order (translation units built with and without _USE_AVX_DEFINES in one program?)");
#else
            check(pl_set_scalar_order(0), "pl_set_scalar_order (translation units built with and without _USE_AVX_DEFINES in one program?)");
#endif
            h = pl_lattice_create(kind, lx, ly, lz, peid, mx, my, mz);
            if (!h) check(1, "pl_lattice_create");
            gen = next_gen();
            live().push_back(gen);
            check(plh_lattice_attach_views(h, f0, f), "plh_lattice_attach_views");
        }
        void destroy() {
            if (!h) return;
            plh_lattice_detach(h);
            for (pl_bc* b : planes) pl_bc_destroy(b);
            for (pl_filter* f : filters) pl_filter_destroy(f);
            pl_lattice_destroy(h);
            h = nullptr;
            for (size_t k = 0; k < live().size(); ++k) if (live()[k] == gen) { live().erase(live().begin() + k); break; }
        }
    };

    // ---- closure identity ---------------------------------------------------------------------------------------------------
    // The reference calls the user's lambdas for every boundary site on every call.  Here a plane is evaluated once and baked
    // into device arrays; whether a later call may reuse it is decided from the bytes of the closure objects.  Captures by value
    // are those bytes.  Captures by REFERENCE ([&], as test/nssens.cpp and test/fsi.cpp use) and captured pointers are addresses:
    // the value they refer to is appended to the key as well (the first bytes behind every pointer-like word that can be read),
    // so a changed inlet velocity or time step is seen.  What this cannot see — state behind two indirections, a global the
    // lambda reads — is caught by re-evaluating each call site every PANSLBM_B200_REVALIDATE-th call (default 256) and comparing
    // the content; a site caught changing that way is evaluated on every call from then on.  plh_bc_invalidate() drops all keys.
    inline bool lattice_alive(unsigned long long gen) { for (unsigned long long g : Core::live()) if (g == gen) return true; return false; }
    template<class F> inline void append_bytes(std::string& s, const F& f) {
        if constexpr (!std::is_empty<F>::value) s.append(reinterpret_cast<const char*>(&f), sizeof(F));
        s.push_back('|');
    }
    // up to `n` bytes at `addr` without ever faulting (a pipe write reports EFAULT instead); memory this library mirrors is skipped:
    // a lambda reading a FIELD is not pure and is handled by revalidation
    inline size_t peek(const void* addr, char* out, size_t n) {
        static int fd[2] = {-1, -1};
        if (fd[0] < 0 && pipe(fd) != 0) return 0;
        if (plh_owns_range(addr)) return 0;
        const ssize_t w = write(fd[1], addr, n);
        if (w <= 0) return 0;
        const ssize_t r = read(fd[0], out, (size_t)w);
        return r > 0 ? (size_t)r : 0;
    }
    template<class F> inline void append_pointees(std::string& s, const F& f) {
        if constexpr (!std::is_empty<F>::value && sizeof(F) >= sizeof(void*)) {
            const char* b = reinterpret_cast<const char*>(&f);
            for (size_t o = 0; o + sizeof(void*) <= sizeof(F); o += sizeof(void*)) {
                unsigned long long v;
                std::memcpy(&v, b + o, sizeof(v));
                if (v < 0x10000ull || v >= 0x0000800000000000ull || (v & 3ull)) continue;      // not a user-space address of an int / double / object
                char buf[32];
                const size_t got = peek(reinterpret_cast<const void*>(v), buf, sizeof(buf));
                s.append(buf, got);
                s.push_back('^');
            }
        }
    }
    inline unsigned long long& bake_epoch() { static unsigned long long e = 0; return e; }      // bumped by invalidate(): every key is stale
    inline void invalidate() { ++bake_epoch(); }
    inline int revalidate_every() {
        static int n = -1;
        if (n < 0) { const char* v = std::getenv("PANSLBM_B200_REVALIDATE"); n = v && *v ? std::atoi(v) : 256; if (n < 1) n = 1; }
        return n;
    }
    template<int ND, class F> inline double call_value(F& f, int i, int j, int k) {
        if constexpr (std::is_same<F, none_t>::value) { (void)f; (void)i; (void)j; (void)k; return 0.0; }
        else if constexpr (ND == 2) { (void)k; return (double)f(i, j); }
        else return (double)f(i, j, k);
    }
    template<int ND, class F> inline int call_mask(F& f, int i, int j, int k) {
        if constexpr (ND == 2) { (void)k; return (int)f(i, j); }
        else return (int)f(i, j, k);
    }

    // One plane closure of lattice `p` (plane axis = GLOBAL coord, outward dir): the callables evaluated on the local plane
    // sites with global coordinates, as the reference does (navierstokes.h:155-157), baked into a pl_bc.  Cached per call
    // site (= per instantiation and plane) by the closure key above — a few variants per site, for code that alternates between
    // closures — and per lattice by content: equal planes share one pl_bc, which keeps the handles stable for the fusion engine.
    // A call site that keeps producing new content (a time-dependent inlet; a closure the key cannot see through) turns VOLATILE:
    // it is evaluated on every call into a plane of its own whose value arrays are replaced in place (pl_bc_update_values), so
    // the handle — and with it the fused plan — survives and no memory accumulates; only a changed MASK makes a new plane.
    template<class P, class Fm, class F0, class F1, class F2>
    const pl_bc* baked(P& p, int type, int axis, int coord, int dir, Fm mask, F0 v0, F1 v1, F2 v2) {
        struct Variant { std::string key; const pl_bc* bc; int plane; };
        struct Site { unsigned long long gen, epoch; int type, axis, coord, dir; std::vector<Variant> var; unsigned calls, misses; int own; bool vol; };
        static std::vector<Site> cache;
        constexpr bool cacheable = std::is_trivially_copyable<Fm>::value && std::is_trivially_copyable<F0>::value &&
                                   std::is_trivially_copyable<F1>::value && std::is_trivially_copyable<F2>::value;
        Core& core = p.b200_core();
        Site* site = nullptr;
        for (Site& e : cache) if (e.gen == core.gen && e.type == type && e.axis == axis && e.coord == coord && e.dir == dir) { site = &e; break; }
        if (!site) {
            if (cache.size() >= 256) cache.erase(cache.begin(), cache.begin() + 128);
            cache.push_back(Site{core.gen, bake_epoch(), type, axis, coord, dir, {}, 0u, 0u, -1, !cacheable});
            site = &cache.back();
        }
        if (site->epoch != bake_epoch()) { site->var.clear(); site->epoch = bake_epoch(); site->misses = 0; }
        std::string ckey;
        Variant* hit = nullptr;
        if constexpr (cacheable) {
            if (!site->vol) {
                append_bytes(ckey, mask); append_bytes(ckey, v0); append_bytes(ckey, v1); append_bytes(ckey, v2);
                append_pointees(ckey, mask); append_pointees(ckey, v0); append_pointees(ckey, v1); append_pointees(ckey, v2);
                for (Variant& v : site->var) if (v.key == ckey) { hit = &v; break; }
                if (hit && (++site->calls % (unsigned)revalidate_every()) != 0) return hit->bc;
            }
        }
        constexpr int ND = P::nd;
        const int off[3] = {p.offsetx, p.offsety, p.offsetz}, n[3] = {p.nx, p.ny, p.nz};
        const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
        const int loc = coord - off[axis];
        const bool local = 0 <= loc && loc < n[axis];
        const size_t np = local ? (size_t)n[a1]*n[a2] : 0;
        std::vector<uint8_t> m(np);
        constexpr bool h0 = !std::is_same<F0, none_t>::value, h1 = !std::is_same<F1, none_t>::value, h2 = !std::is_same<F2, none_t>::value;
        std::vector<double> a(h0 ? np : 0), b(h1 ? np : 0), c(h2 ? np : 0);
        const bool raw = type == PL_BC_BOUNCE || type == PL_BC_IBOUNCE || type == PL_BC_AAD_ISET_RHO;
        for (int s2 = 0; s2 < (local ? n[a2] : 0); ++s2) {
            for (int s1 = 0; s1 < n[a1]; ++s1) {
                int g[3];
                g[axis] = coord; g[a1] = s1 + off[a1]; g[a2] = s2 + off[a2];
                const size_t t = (size_t)s1 + (size_t)n[a1]*s2;
                const int mv = call_mask<ND>(mask, g[0], g[1], g[2]);
                m[t] = raw ? (uint8_t)((mv == 1 || mv == 2) ? mv : 0) : (uint8_t)(mv != 0);
                if constexpr (h0) a[t] = call_value<ND>(v0, g[0], g[1], g[2]);
                if constexpr (h1) b[t] = call_value<ND>(v1, g[0], g[1], g[2]);
                if constexpr (h2) c[t] = call_value<ND>(v2, g[0], g[1], g[2]);
            }
        }
        // content key: header + mask, then the values
        std::string head;
        head.append(reinterpret_cast<const char*>(&type), sizeof(int)); head.append(reinterpret_cast<const char*>(&axis), sizeof(int));
        head.append(reinterpret_cast<const char*>(&coord), sizeof(int)); head.append(reinterpret_cast<const char*>(&dir), sizeof(int));
        head.append(reinterpret_cast<const char*>(m.data()), m.size());
        std::string key = head;
        key.append(reinterpret_cast<const char*>(a.data()), a.size()*sizeof(double)); key.push_back('|');
        key.append(reinterpret_cast<const char*>(b.data()), b.size()*sizeof(double)); key.push_back('|');
        key.append(reinterpret_cast<const char*>(c.data()), c.size()*sizeof(double));
        auto create = [&](bool exclusive) {
            pl_bc* nb = pl_bc_create(core.h, type, axis, coord, dir, local ? m.data() : nullptr, h0 && local ? a.data() : nullptr,
                                     h1 && local ? b.data() : nullptr, h2 && local ? c.data() : nullptr);
            if (!nb) check(1, "pl_bc_create");
            core.planes.push_back(nb); core.contents.push_back(key); core.exclusive.push_back(exclusive ? 1 : 0);
            return (int)core.planes.size() - 1;
        };
        if (hit) {
            if (core.contents[hit->plane] == key) return hit->bc;      // revalidated: the closure still produces what was baked
            site->vol = true;                                           // same closure bytes, other content: not a pure closure
        }
        if (!site->vol && site->misses >= 4) site->vol = true;          // keeps changing: stop making planes
        if (!site->vol) {
            ++site->misses;
            int plane = -1;
            for (size_t k = 0; k < core.contents.size(); ++k) if (!core.exclusive[k] && core.contents[k] == key) { plane = (int)k; break; }
            if (plane < 0) plane = create(false);
            if (site->var.size() >= 4) site->var.erase(site->var.begin());
            site->var.push_back(Variant{ckey, core.planes[plane], plane});
            return core.planes[plane];
        }
        // volatile site: its own plane, values replaced in place while the mask stays
        if (site->own >= 0 && core.contents[site->own].size() == key.size() && core.contents[site->own].compare(0, head.size(), head) == 0) {
            if (core.contents[site->own] != key) {
                check(plh_bc_update_values(core.planes[site->own], h0 && local ? a.data() : nullptr, h1 && local ? b.data() : nullptr, h2 && local ? c.data() : nullptr),
                      "plh_bc_update_values");
                core.contents[site->own] = key;
            }
        } else {
            site->own = create(true);
        }
        return core.planes[site->own];
    }

    // The weights of a cone filter of radius _R on lattice `p` (densityfilter.h:389-497): the reference evaluates
    // `_weight(i1,j1,k1,i2,j2,k2)` for every pair within _R on EVERY call; here it is evaluated once per (lattice, _R, callable),
    // site by site, and folded into PATTERNS on the fly: sites with the same (2nR+1)^nd weights share one entry (the drivers'
    // weights depend on the offset and on which side of the design box the two sites lie, production/heatsink3D.cpp:87-93: a
    // few hundred patterns), so neither the host nor the device ever holds a per-site table.  Pairs beyond _R or outside the
    // domain get weight 0 (they do not enter the reference's sums).
    template<class P, class F>
    pl_filter* filter(P& p, double _R, F _weight) {
        struct Entry { unsigned long long gen, epoch; double R; std::string bytes; pl_filter* f; };
        static std::vector<Entry> cache;
        Core& core = p.b200_core();
        std::string bytes;
        constexpr bool cacheable = std::is_trivially_copyable<F>::value;
        if constexpr (cacheable) {
            append_bytes(bytes, _weight); append_pointees(bytes, _weight);
            for (const Entry& e : cache) if (e.gen == core.gen && e.epoch == bake_epoch() && e.R == _R && e.bytes == bytes) return e.f;
        }
        // entries of lattices that are gone (their filters were destroyed with them) leave the cache
        for (size_t k = 0; k < cache.size();) { if (cache[k].gen != core.gen && !lattice_alive(cache[k].gen)) cache.erase(cache.begin() + k); else ++k; }
        const int nR = (int)_R, side = 2*nR + 1;
        const size_t n = (size_t)p.nxyz, K = (size_t)side*side*side;
        std::vector<double> patterns, row(K);
        std::vector<int> pid(n);
        std::unordered_map<std::string, int> seen;
        for (int k1 = 0; k1 < p.nz; ++k1)
            for (int j1 = 0; j1 < p.ny; ++j1)
                for (int i1 = 0; i1 < p.nx; ++i1) {
                    size_t o = 0;
                    for (int i2 = i1 - nR; i2 <= i1 + nR; ++i2)
                        for (int j2 = j1 - nR; j2 <= j1 + nR; ++j2)
                            for (int k2 = k1 - nR; k2 <= k1 + nR; ++k2, ++o) {
                                row[o] = 0.0;
                                if (P::nd == 2 && k2 != k1) continue;
                                // neighbours anywhere in the GLOBAL domain count (own block or another rank's: heavisidefilter.h:470-556)
                                if (i2 + p.offsetx < 0 || i2 + p.offsetx >= p.lx || j2 + p.offsety < 0 || j2 + p.offsety >= p.ly ||
                                    k2 + p.offsetz < 0 || k2 + p.offsetz >= p.lz) continue;
                                const double distance = std::sqrt(std::pow(i1 - i2, 2.0) + std::pow(j1 - j2, 2.0) + std::pow(k1 - k2, 2.0));
                                if (distance <= _R)
                                    row[o] = (double)_weight(i1 + p.offsetx, j1 + p.offsety, k1 + p.offsetz, i2 + p.offsetx, j2 + p.offsety, k2 + p.offsetz);
                            }
                    auto r = seen.emplace(std::string(reinterpret_cast<const char*>(row.data()), K*sizeof(double)), (int)seen.size());
                    if (r.second) patterns.insert(patterns.end(), row.begin(), row.end());
                    pid[(size_t)p.Index(i1, j1, k1)] = r.first->second;
                }
        pl_filter* f = pl_filter_create_patterns(core.h, nR, patterns.data(), (int)seen.size(), pid.data());
        if (!f) check(1, "pl_filter_create_patterns");
        core.filters.push_back(f);
        if constexpr (cacheable) cache.push_back(Entry{core.gen, bake_epoch(), _R, bytes, f});
        return f;
    }

    inline pl_bc_aux aux(const double* rho, const double* ux, const double* uy, const double* uz, const double* tem, const double* kfield, double kconst, double eps) {
        pl_bc_aux a;
        std::memset(&a, 0, sizeof(a));
        a.rho = rho; a.ux = ux; a.uy = uy; a.uz = uz; a.tem = tem; a.diffusivity = kfield; a.diffusivity_const = kconst; a.eps = eps;
        return a;
    }
    // apply one closure type on one plane
    template<class P, class Fm, class F0, class F1, class F2>
    inline void plane(P& p, int type, int axis, int coord, int dir, Fm mask, F0 v0, F1 v1, F2 v2, const pl_bc_aux* a, pl_lattice* other = nullptr) {
        const pl_bc* bc = baked(p, type, axis, coord, dir, mask, v0, v1, v2);
        check(plh_bc(p.b200_core().h, other, bc, a), "plh_bc");
    }
    // ... and on the 2*nd faces of the global domain in the reference's order xmin,xmax,ymin,ymax[,zmin,zmax] (d3q15.h:182-189)
    template<class P, class Fm, class F0, class F1, class F2>
    inline void faces(P& p, int type, Fm mask, F0 v0, F1 v1, F2 v2, const pl_bc_aux* a, pl_lattice* other = nullptr) {
        const int ext[3] = {p.lx, p.ly, p.lz};
        for (int axis = 0; axis < P::nd; ++axis) {
            plane(p, type, axis, 0, -1, mask, v0, v1, v2, a, other);
            plane(p, type, axis, ext[axis] - 1, 1, mask, v0, v1, v2, a, other);
        }
    }

    inline pl_collide_args collide_args(int model, bool issave, double viscosity) {
        pl_collide_args a;
        std::memset(&a, 0, sizeof(a));
        a.model = model; a.issave = issave ? 1 : 0; a.viscosity = viscosity;
        return a;
    }
    template<class T> inline void only_double() { static_assert(std::is_same<T, double>::value, "panslbm_b200 computes in fp64: instantiate with T = double"); }
}  // namespace b200
}  // namespace PANSLBM2
