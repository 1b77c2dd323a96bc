// Render-side types the reference's physics sources mention (scene cloning, vehicle.cpp's meshes). A PHYSICS_ONLY build has no
// renderer: these are inert stand-ins with just the members those sources touch. Test infrastructure only.
#pragma once
#include "pch.h"
#include "core/math.h"
struct submesh_info { int unused = 0; };
struct pbr_material { int unused = 0; };
struct dx_mesh { int unused = 0; };
struct submesh { submesh_info info; int aabb; trs transform; ref<pbr_material> material; };
struct multi_mesh { std::vector<submesh> submeshes; dx_mesh mesh; };
struct mesh_component { ref<multi_mesh> mesh; mesh_component() {} mesh_component(ref<multi_mesh> m) : mesh(m) {} };
struct raytrace_component { int unused = 0; };      // (not empty: EnTT's empty-type optimisation makes emplace<> return void,
struct animation_component { int unused = 0; };     //  which scene.h's addComponent binds to a reference)
struct box_mesh_desc { vec3 center = vec3(0.f); vec3 radius = vec3(1.f); quat rotation = quat::identity; };
struct cylinder_mesh_desc { vec3 center = vec3(0.f); float height = 1.f, radius = 1.f; uint32 slices = 16; quat rotation = quat::identity; };
struct hollow_cylinder_mesh_desc { vec3 center = vec3(0.f); float height = 1.f, radius = 1.f, innerRadius = 0.5f; uint32 slices = 16; quat rotation = quat::identity; };
struct capsule_mesh_desc { vec3 center = vec3(0.f); float height = 1.f, radius = 1.f; quat rotation = quat::identity; };
struct mesh_builder
{
    void pushBox(const box_mesh_desc&) {}
    void pushCylinder(const cylinder_mesh_desc&) {}
    void pushHollowCylinder(const hollow_cylinder_mesh_desc&) {}
    void pushCapsule(const capsule_mesh_desc&) {}
    submesh_info endSubmesh() { return submesh_info(); }
    dx_mesh createDXMesh() { return dx_mesh(); }
};
static inline ref<pbr_material> createPBRMaterial(std::initializer_list<const char*>) { return make_ref<pbr_material>(); }
