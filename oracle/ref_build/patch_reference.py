#!/usr/bin/env python
"""Makes the reference's OWN physics sources compile with g++ -- TEST INFRASTRUCTURE (oracle/_ref), never the product.

The reference (pkurth/D3D12Renderer, /root/reference) is MSVC-only. Its "Physics-Lib" target (premake5.lua:389-465, PHYSICS_ONLY)
is a self-contained subset: src/physics, src/core/math, src/scene (EnTT), src/learning. This script copies exactly those files
from the read-only reference tree into oracle/_ref/src (git-ignored, never committed) and applies the SMALLEST set of textual
rewrites g++ 13 needs; every rule below names the MSVC-ism it removes. No arithmetic is touched: expressions, operation order,
constants and control flow of the physics code stay byte-for-byte what the reference has.

usage: patch_reference.py <reference root> <output dir>
"""
import os
import re
import sys

FILES = """
physics/bounding_volumes.h physics/bounding_volumes.cpp physics/bounding_volumes_simd.h
physics/collision_broad.h physics/collision_broad.cpp physics/collision_epa.h physics/collision_epa.cpp
physics/collision_gjk.h physics/collision_gjk.cpp physics/collision_narrow.h physics/collision_narrow.cpp physics/collision_sat.h
physics/constraints.h physics/constraints.cpp physics/physics.h physics/physics.cpp physics/rigid_body.h physics/rigid_body.cpp
physics/ragdoll.h physics/ragdoll.cpp physics/vehicle.h physics/vehicle.cpp physics/cloth.h physics/cloth.cpp
physics/heightmap_collision.h physics/heightmap_collision.cpp terrain/heightmap_collider.h terrain/heightmap_collider.cpp
core/math.h core/math.cpp core/simd.h core/math_simd.h core/soa.h core/memory.h core/random.h core/reflect.h core/preprocessor_for_each.h
scene/scene.h scene/scene.cpp scene/components.h learning/learned_locomotion.h learning/learned_locomotion.cpp
""".split()


def rewrite_union(text, name, drop_if, add_members):
    """Inside `union <name> { ... };` remove the anonymous structs whose body matches drop_if (members with constructors are not
    allowed in anonymous aggregates in ISO C++ / g++; MSVC accepts them) and add the aliases the code uses as direct union members."""
    m = re.search(r"union\s+%s\s*\{" % name, text)
    assert m, name
    start = m.end()
    depth, i = 1, start
    while depth:
        c = text[i]
        depth += (c == "{") - (c == "}")
        i += 1
    body = text[start:i - 1]
    body = re.sub(r"struct\s*\{([^{}]*)\};", lambda s: "" if re.search(drop_if, s.group(1)) else s.group(0), body)
    return text[:start] + "\n\t" + " ".join(add_members) + body + text[i - 1:]


def patch_math_h(t):
    # MSVC extension: anonymous structs inside unions holding members with constructors, and the same member name in two of them
    t = rewrite_union(t, "vec3", r"\bvec2\b", ["vec2 xy;"])
    t = rewrite_union(t, "vec4", r"\bvec2\b|\bvec3\b", ["vec3 xyz;", "vec2 xy;"])
    t = rewrite_union(t, "quat", r"\bvec3\b", ["vec3 v;", "struct { float _x, _y, _z, cosHalfAngle; };"])
    t = rewrite_union(t, "mat2", r"\bvec2\b", [])
    t = rewrite_union(t, "mat3", r"\bvec3\b", [])
    t = rewrite_union(t, "mat4", r"\bvec4\b|\bw4_float\b", ["w4_float f4cols[4];"])
    return t


def union_to_struct(text, name, elem, arr, count):
    """`union <name> { struct { <elem> a, b, ...; }; <elem> <arr>[<count>]; ... }` (named views of an array; MSVC allows constructors
    in the anonymous struct) -> a struct with the named members and a zero-size indexable view over them (shim pch.h: ref_array_view)."""
    m = re.search(r"union\s+%s\s*\{" % name, text)
    assert m, name
    start = m.end()
    depth, i = 1, start
    while depth:
        c = text[i]
        depth += (c == "{") - (c == "}")
        i += 1
    body = text[start:i - 1]
    s = re.search(r"struct\s*\{([^{}]*)\};", body)
    assert s, name
    body = body[:s.start()] + s.group(1) + body[s.end():]
    body, n = re.subn(r"\b%s\s+%s\s*\[\s*%d\s*\]\s*;" % (elem, arr, count), "ref_array_view<%s, %d> %s;" % (elem, count, arr), body)
    assert n == 1, (name, n)
    return text[:m.start()] + "struct " + name + "\n{" + body + text[i - 1:]


def wide_union_to_struct(text, name, aliases=()):
    """math_simd.h: `template <typename simd_t> union wN_x { struct { simd_t x, y, ..; }; [alias structs]; simd_t data[N]; ... }`.
    simd_t (w4_float / w8_float) has constructors, so g++ rejects the anonymous struct on instantiation. -> struct with the named
    lanes, a zero-size indexable view in front (ref_array_front, [[no_unique_address]]: the size stays N * sizeof(simd_t)) and, for
    wN_quat, reference members for the .v / .v4 aliases (copy operations written out so that the references never rebind)."""
    m = re.search(r"union\s+%s\s*\{" % name, text)
    assert m, name
    start = m.end()
    depth, i = 1, start
    while depth:
        c = text[i]
        depth += (c == "{") - (c == "}")
        i += 1
    body = text[start:i - 1]
    structs = list(re.finditer(r"struct\s*\{([^{}]*)\};", body))
    first = structs[0]
    names = [n.strip() for n in re.sub(r"\bsimd_t\b", "", first.group(1)).replace(";", ",").split(",") if n.strip()]
    out = body
    for st in reversed(structs[1:]):
        out = out[:st.start()] + out[st.end():]
    out = out.replace(first.group(0), "simd_t " + ", ".join(names) + ";", 1)
    out, n = re.subn(r"\bsimd_t\s+(\w+)\s*\[\s*\d+\s*\]\s*;", "", out)
    arr = re.search(r"\bsimd_t\s+(\w+)\s*\[\s*\d+\s*\]\s*;", body).group(1)
    assert n == 1, (name, n)
    extra = ""
    for (typ, alias) in aliases:
        out = re.sub(r"\b%s\s+%s\s*;" % (re.escape(typ), alias), "", out)
        extra += "\n\t%s& %s = *reinterpret_cast<%s*>(&%s);" % (typ, alias, typ, names[0])
    if aliases:
        extra += "\n\t%s(const %s& o) : %s {}" % (name, name, ", ".join("%s(o.%s)" % (k, k) for k in names))
        extra += "\n\t%s& operator=(const %s& o) { %s return *this; }" % (name, name, " ".join("%s = o.%s;" % (k, k) for k in names))
    head = "\n\t[[no_unique_address]] ref_array_front<simd_t> %s;" % arr
    return text[:m.start()] + "struct " + name + "\n{" + head + out + extra + "\n" + text[i - 1:]


def patch_math_simd_h(t):
    for name in ("wN_vec2", "wN_vec3", "wN_vec4", "wN_mat2", "wN_mat3", "wN_mat4"):
        t = wide_union_to_struct(t, name)
    t = wide_union_to_struct(t, "wN_quat", aliases=(("wN_vec3<simd_t>", "v"), ("wN_vec4<simd_t>", "v4")))
    return t


def anonymous_unions_to_views(t):
    """`union { struct { T a; T b; ... }; T arr[N]; };` as an anonymous member (ragdoll.h, vehicle.h: named entities + array) with a T
    that has constructors -> the named members followed by the indexable view (see union_to_struct)."""
    def repl(m):
        return m.group(1) + "\n\tref_array_view<%s, %s> %s;" % (m.group(2), m.group(4), m.group(3))
    return re.sub(r"union\s*\{\s*struct\s*\{([^{}]*)\};\s*(\w+)\s+(\w+)\s*\[\s*(\d+)\s*\]\s*;\s*\};", repl, t)


def generic(t):
    # MSVC accepts a storage class on explicit specialisations
    t = re.sub(r"template\s*<>\s*static\b", "template <>", t)
    # MSVC-only named lanes of the vector types (__m128::m128_f32 ...): index through a pointer instead
    t = re.sub(r"((?:this->)?\w+)\.m(?:128|256|512)_f32\[", r"((float*)&(\1))[", t)
    t = re.sub(r"((?:this->)?\w+)\.m(?:128|256|512)i_i32\[", r"((int*)&(\1))[", t)
    # aliases removed from the unions above
    t = re.sub(r"\.col([0-3])\b", r".cols[\1]", t)
    t = re.sub(r"\.f4([0-3])\b", r".f4cols[\1]", t)
    return t


def patch(rel, t):
    t = generic(t)
    if rel == "core/math.h":
        t = patch_math_h(t)
    if rel == "physics/bounding_volumes.h":
        t = union_to_struct(t, "bounding_box_corners", "vec3", "corners", 8)
    if rel == "physics/collision_gjk.h":
        t = union_to_struct(t, "extruded_triangle_support_fn", "vec3", "points", 6)
    if rel in ("physics/bounding_volumes.cpp", "physics/collision_narrow.cpp"):
        t = union_to_struct(t, "obb_axes", "vec3", "u", 3)
    if rel == "core/math_simd.h":
        t = patch_math_simd_h(t)
        t = re.sub(r"\bwN_quat result;", "wN_quat<simd_t> result;", t)     # class template argument deduction from nothing (MSVC: lazy template bodies)
    if rel == "physics/bounding_volumes_simd.h":
        t = t.replace("wN_bounding_box wN_bounding_box::from", "wN_bounding_box from")     # extra qualification on in-class members
        t = re.sub(r"\bw_vec3\b", "wN_vec3<simd_t>", t); t = re.sub(r"\bw_float\b", "simd_t", t)   # names of a LATER typedef inside a template
    if rel == "physics/collision_narrow.cpp":
        t = re.sub(r"auto&\s*\[([^\]]*)\]\s*=\s*writeContext\.pushContact\(\)", r"auto&& [\1] = writeContext.pushContact()", t)   # lvalue ref to a temporary pair
        t = re.sub(r"=\s*(scalar_to_wide<\w+>::type)", r"= typename \1", t)                                                      # missing typename
    if rel in ("physics/ragdoll.h", "physics/vehicle.h"):
        t = anonymous_unions_to_views(t)
    if rel == "scene/scene.h":
        # `struct x_component` inside the addComponent member template names a LATER namespace-scope type; without a prior declaration
        # ISO C++ makes it a new local (incomplete) type. Declare them up front.
        names = sorted(set(re.findall(r"struct\s+(\w+_component)\b", t)))
        t = t.replace("struct game_scene;", "struct game_scene;\n" + "".join("struct %s;\n" % n for n in names), 1)
    if rel == "physics/physics.cpp":
        # give the test harness (ref_api.cpp) access to two file-local items: the hull geometry list and getWorldSpaceColliders
        t = t.replace("static std::vector<bounding_hull_geometry> boundingHullGeometries;", "std::vector<bounding_hull_geometry> boundingHullGeometries;", 1)
        t = t.replace("static void getWorldSpaceColliders(", "void getWorldSpaceColliders(", 1)
    if rel == "learning/learned_locomotion.cpp":
        # the environment seeds its push generator with the wall clock and hard-codes the push probability: make both settable by the harness
        t = t.replace("static random_number_generator rng = { (uint32)time(0) };", "random_number_generator rng = { 1 }; float refPushProbability = 0.02f;", 1)
        t = t.replace("rng.randomFloat01() < 0.02f", "rng.randomFloat01() < refPushProbability", 1)
        # ... and let it choose between the scalar path (the parity path) and the AVX2 path the DLL uses by default (physics.h:394-396)
        t = t.replace("float refPushProbability = 0.02f;", "float refPushProbability = 0.02f; bool refEnvSimd = true;", 1)
        t = t.replace("physicsSettings.frameRate = 60;", "physicsSettings.frameRate = 60; physicsSettings.simdBroadPhase = physicsSettings.simdNarrowPhase = physicsSettings.simdConstraintSolver = refEnvSimd;", 1)
    if rel == "core/random.h":
        t = re.sub(r"\bn\.yzw\b", "vec3(n.y, n.z, n.w)", t)
        t = re.sub(r"\bn\.yz\b", "vec2(n.y, n.z)", t)
    return t


def main():
    ref, out = sys.argv[1], sys.argv[2]
    for rel in FILES:
        src = os.path.join(ref, "src", rel)
        dst = os.path.join(out, "src", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, encoding="utf-8", errors="replace") as f:
            t = f.read()
        with open(dst, "w", encoding="utf-8") as f:
            f.write(patch(rel, t))
    print("patched %d reference files into %s" % (len(FILES), os.path.join(out, "src")))


if __name__ == "__main__":
    main()
