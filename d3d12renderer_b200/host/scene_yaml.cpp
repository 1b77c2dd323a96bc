// scene_yaml.cpp -- reads / writes the physics components of the reference's scene files (src/scene/serialization_yaml.cpp:74-240,
// 395-430, 470-520; math types src/core/yaml.h): per entity "Tag", "Transform" {Position, Rotation, Scale}, "Rigid body"
// {Local COG, Inv mass, Inv inertia, Gravity factor, Linear damping, Angular damping}, "Force field" {Force}, "Colliders" [ {Type,
// shape keys, Restitution, Friction, Density} ]. Everything else in a reference scene (camera, lights, meshes, cloth, terrain) is
// skipped on load and not written on save. The reference uses yaml-cpp; this is a small reader for the block / flow subset that
// yaml-cpp emits for these files (block maps and sequences by indentation, flow sequences of scalars, plain scalars).
#include "physics_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace b200host {

namespace {

struct YNode
{
    enum Kind { Null, Scalar, Map, Seq } kind = Null;
    std::string scalar;
    std::vector<std::pair<std::string, YNode>> map;
    std::vector<YNode> seq;
    const YNode* find(const char* key) const { for (const auto& kv : map) if (kv.first == key) return &kv.second; return nullptr; }
};

struct Line { int indent; std::string text; };

std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

YNode parseFlowOrScalar(const std::string& v)
{
    YNode n;
    const std::string t = trim(v);
    if (!t.empty() && t.front() == '[' && t.back() == ']')
    {
        n.kind = YNode::Seq;
        std::string inner = t.substr(1, t.size() - 2), item;
        std::stringstream ss(inner);
        while (std::getline(ss, item, ',')) { YNode c; c.kind = YNode::Scalar; c.scalar = trim(item); if (!c.scalar.empty()) n.seq.push_back(c); }
        return n;
    }
    n.kind = t.empty() ? YNode::Null : YNode::Scalar;
    n.scalar = t;
    if (n.scalar.size() >= 2 && ((n.scalar.front() == '"' && n.scalar.back() == '"') || (n.scalar.front() == '\'' && n.scalar.back() == '\''))) n.scalar = n.scalar.substr(1, n.scalar.size() - 2);
    return n;
}

// splits "key: value" at the first ": " (or trailing ':') outside quotes / brackets
bool splitKey(const std::string& t, std::string& key, std::string& value)
{
    int depth = 0; bool quoted = false;
    for (size_t i = 0; i < t.size(); ++i)
    {
        const char ch = t[i];
        if (ch == '"') quoted = !quoted;
        if (quoted) continue;
        if (ch == '[' || ch == '{') ++depth; else if (ch == ']' || ch == '}') --depth;
        if (ch == ':' && depth == 0 && (i + 1 == t.size() || t[i + 1] == ' ')) { key = trim(t.substr(0, i)); value = i + 1 < t.size() ? trim(t.substr(i + 1)) : std::string(); return true; }
    }
    return false;
}

YNode parseBlock(const std::vector<Line>& lines, size_t& i, int indent);

// a sequence item "- something" at `indent`: the item body is a scalar, or a map whose first key sits on the dash line
YNode parseSeq(const std::vector<Line>& lines, size_t& i, int indent)
{
    YNode n; n.kind = YNode::Seq;
    while (i < lines.size() && lines[i].indent == indent && lines[i].text.rfind("-", 0) == 0 && (lines[i].text.size() == 1 || lines[i].text[1] == ' '))
    {
        const std::string rest = trim(lines[i].text.substr(1));
        const int childIndent = indent + 2;
        std::string k, v;
        if (rest.empty()) { ++i; n.seq.push_back(parseBlock(lines, i, i < lines.size() ? lines[i].indent : childIndent)); }
        else if (splitKey(rest, k, v))
        {   // re-interpret the dash line as the first line of a map indented by two
            std::vector<Line> tmp; tmp.push_back(Line{ childIndent, rest });
            size_t j = i + 1;
            while (j < lines.size() && lines[j].indent >= childIndent) { tmp.push_back(lines[j]); ++j; }
            size_t t = 0; n.seq.push_back(parseBlock(tmp, t, childIndent));
            i = j;
        }
        else { n.seq.push_back(parseFlowOrScalar(rest)); ++i; }
    }
    return n;
}

YNode parseBlock(const std::vector<Line>& lines, size_t& i, int indent)
{
    if (i >= lines.size()) return YNode();
    if (lines[i].text.rfind("-", 0) == 0 && (lines[i].text.size() == 1 || lines[i].text[1] == ' ')) return parseSeq(lines, i, lines[i].indent);
    YNode n; n.kind = YNode::Map;
    while (i < lines.size() && lines[i].indent == indent)
    {
        std::string k, v;
        if (!splitKey(lines[i].text, k, v)) { ++i; continue; }
        ++i;
        if (!v.empty()) n.map.emplace_back(k, parseFlowOrScalar(v));
        else if (i < lines.size() && (lines[i].indent > indent || (lines[i].indent == indent && lines[i].text.rfind("- ", 0) == 0)))
            n.map.emplace_back(k, parseBlock(lines, i, lines[i].indent));     // yaml-cpp emits sequences of a map value at the key's own indent or deeper
        else n.map.emplace_back(k, YNode());
    }
    return n;
}

YNode parseYaml(std::istream& in)
{
    std::vector<Line> lines; std::string raw;
    while (std::getline(in, raw))
    {
        const size_t hash = raw.find(" #");
        if (hash != std::string::npos) raw = raw.substr(0, hash);
        const std::string t = trim(raw);
        if (t.empty() || t[0] == '#' || t == "---") continue;
        int indent = 0; while ((size_t)indent < raw.size() && raw[indent] == ' ') ++indent;
        lines.push_back(Line{ indent, t });
    }
    size_t i = 0;
    return lines.empty() ? YNode() : parseBlock(lines, i, lines[0].indent);
}

bool getFloats(const YNode* n, float* out, size_t count)
{
    if (!n || n->kind != YNode::Seq || n->seq.size() != count) return false;
    for (size_t i = 0; i < count; ++i) out[i] = strtof(n->seq[i].scalar.c_str(), nullptr);
    return true;
}
bool getFloat(const YNode* n, float& out) { if (!n || n->kind != YNode::Scalar) return false; out = strtof(n->scalar.c_str(), nullptr); return true; }
bool getVec3(const YNode* n, vec3& v) { float f[3]; if (!getFloats(n, f, 3)) return false; v = b200::mk3(f[0], f[1], f[2]); return true; }
bool getQuat(const YNode* n, quat& q) { float f[4]; if (!getFloats(n, f, 4)) return false; q = b200::mkq(f[0], f[1], f[2], f[3]); return true; }

const char* kTypeNames[6] = { "Sphere", "Capsule", "Cylinder", "AABB", "OBB", "Hull" };    // colliderTypeNames, physics.h:72-80

void putF(std::ostream& o, float v) { char buf[48]; snprintf(buf, sizeof(buf), "%.9g", (double)v); o << buf; }
void putVec3(std::ostream& o, vec3 v) { o << '['; putF(o, v.x); o << ", "; putF(o, v.y); o << ", "; putF(o, v.z); o << ']'; }
void putQuat(std::ostream& o, quat q) { o << '['; putF(o, q.x); o << ", "; putF(o, q.y); o << ", "; putF(o, q.z); o << ", "; putF(o, q.w); o << ']'; }

} // namespace

// serialization_yaml.cpp:395-430 restricted to the physics components
bool saveSceneYaml(const game_scene& s, std::ostream& o, const char* sceneName)
{
    o << "Scene: " << (sceneName ? sceneName : "Scene") << "\nEntities:\n";
    for (size_t e = 0; e < s.entities.size(); ++e)
    {
        const game_scene::entity_data& d = s.entities[e];
        o << "  - Tag: " << d.name << "\n";
        if (d.hasTransform)
        {
            const trs& t = d.rigidBody >= 0 ? s.physicsTransform1[d.rigidBody] : d.transform;
            o << "    Transform:\n      Position: "; putVec3(o, t.position); o << "\n      Rotation: "; putQuat(o, t.rotation); o << "\n      Scale: "; putVec3(o, t.scale); o << "\n";
        }
        if (d.rigidBody >= 0)
        {
            const rigid_body_component& rb = s.rigidBodies[d.rigidBody];
            o << "    Rigid body:\n      Local COG: "; putVec3(o, rb.localCOGPosition);
            o << "\n      Inv mass: "; putF(o, rb.invMass);
            o << "\n      Inv inertia: [";     // core/yaml.h convert<mat3>: m00 m01 m02 m10 ... (row by row of a column-major matrix)
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { if (r || c) o << ", "; putF(o, rb.invInertia.m[c * 3 + r]); }
            o << "]\n      Gravity factor: "; putF(o, rb.gravityFactor);
            o << "\n      Linear damping: "; putF(o, rb.linearDamping);
            o << "\n      Angular damping: "; putF(o, rb.angularDamping); o << "\n";
        }
        if (d.forceField >= 0) { o << "    Force field:\n      Force: "; putVec3(o, s.forceFields[d.forceField].force); o << "\n"; }
        if (d.numColliders)
        {
            o << "    Colliders:\n";
            for (entity_handle c = d.firstColliderEntity; c != null_entity; c = s.colliders[c].nextEntity)
            {
                const collider_component& col = s.colliders[c];
                o << "      - Type: " << kTypeNames[col.type] << "\n";
                switch (col.type)
                {
                case collider_type_sphere: o << "        Center: "; putVec3(o, col.sphere.center); o << "\n        Radius: "; putF(o, col.sphere.radius); o << "\n"; break;
                case collider_type_capsule: o << "        Position A: "; putVec3(o, col.capsule.positionA); o << "\n        Position B: "; putVec3(o, col.capsule.positionB); o << "\n        Radius: "; putF(o, col.capsule.radius); o << "\n"; break;
                case collider_type_aabb: o << "        Min corner: "; putVec3(o, col.aabb.minCorner); o << "\n        Max corner: "; putVec3(o, col.aabb.maxCorner); o << "\n"; break;
                case collider_type_obb: o << "        Center: "; putVec3(o, col.obb.center); o << "\n        Radius: "; putVec3(o, col.obb.radius); o << "\n        Rotation: "; putQuat(o, col.obb.rotation); o << "\n"; break;
                default: break;       // cylinders and hulls carry no shape keys in the reference's files (serialization_yaml.cpp:128-163)
                }
                o << "        Restitution: "; putF(o, col.material.restitution); o << "\n        Friction: "; putF(o, col.material.friction); o << "\n        Density: "; putF(o, col.material.density); o << "\n";
            }
        }
    }
    return (bool)o;
}

// serialization_yaml.cpp:470-520 restricted to the physics components; returns the number of entities created (-1: not a scene file)
int loadSceneYaml(game_scene& s, std::istream& in)
{
    const YNode root = parseYaml(in);
    const YNode* entities = root.kind == YNode::Map ? root.find("Entities") : nullptr;
    if (!entities || entities->kind != YNode::Seq) return -1;
    int created = 0;
    for (const YNode& en : entities->seq)
    {
        if (en.kind != YNode::Map) continue;
        const YNode* tag = en.find("Tag");
        scene_entity e = s.createEntity(tag && tag->kind == YNode::Scalar ? tag->scalar.c_str() : "Entity");
        ++created;
        if (const YNode* t = en.find("Transform"))
        {
            vec3 p = b200::mk3(0.f), sc = b200::mk3(1.f); quat q = b200::qIdentity();
            getVec3(t->find("Position"), p); getQuat(t->find("Rotation"), q); getVec3(t->find("Scale"), sc);
            e.addTransform(p, q, sc);
        }
        const YNode* rbn = en.find("Rigid body");
        if (rbn)
        {   // LOAD_COMPONENT(rigid_body_component, "Rigid body"): the stored mass properties, velocities start at zero
            float invMass = 1.f, g = 1.f, ld = 0.4f, ad = 0.4f;
            getFloat(rbn->find("Inv mass"), invMass); getFloat(rbn->find("Gravity factor"), g); getFloat(rbn->find("Linear damping"), ld); getFloat(rbn->find("Angular damping"), ad);
            e.addRigidBody(invMass == 0.f, g, ld, ad);
            rigid_body_component& rb = e.rigidBody();
            rb.invMass = invMass;
            getVec3(rbn->find("Local COG"), rb.localCOGPosition);
            float m[9];
            if (getFloats(rbn->find("Inv inertia"), m, 9)) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) rb.invInertia.m[c * 3 + r] = m[r * 3 + c];
        }
        if (const YNode* ff = en.find("Force field")) { vec3 f = b200::mk3(0.f); getVec3(ff->find("Force"), f); e.addForceField(f); }
        if (const YNode* cols = en.find("Colliders"))
        {
            if (cols->kind == YNode::Seq)
                for (const YNode& cn : cols->seq)
                {
                    if (cn.kind != YNode::Map) continue;
                    const YNode* type = cn.find("Type");
                    int ti = -1;
                    for (int k = 0; k < 6; ++k) if (type && type->scalar == kTypeNames[k]) ti = k;
                    physics_material mat = {}; mat.restitution = mat.friction = mat.density = 0.f;
                    getFloat(cn.find("Restitution"), mat.restitution); getFloat(cn.find("Friction"), mat.friction); getFloat(cn.find("Density"), mat.density);
                    vec3 a = b200::mk3(0.f), b = b200::mk3(0.f); quat q = b200::qIdentity(); float r = 0.f;
                    switch (ti)
                    {
                    case collider_type_sphere: getVec3(cn.find("Center"), a); getFloat(cn.find("Radius"), r); e.addCollider(collider_component::asSphere({ a, r }, mat)); break;
                    case collider_type_capsule: getVec3(cn.find("Position A"), a); getVec3(cn.find("Position B"), b); getFloat(cn.find("Radius"), r); e.addCollider(collider_component::asCapsule({ a, b, r }, mat)); break;
                    case collider_type_aabb: getVec3(cn.find("Min corner"), a); getVec3(cn.find("Max corner"), b); e.addCollider(collider_component::asAABB(bounding_box::fromMinMax(a, b), mat)); break;
                    case collider_type_obb: getVec3(cn.find("Center"), a); getVec3(cn.find("Radius"), b); getQuat(cn.find("Rotation"), q); e.addCollider(collider_component::asOBB({ q, a, b }, mat)); break;
                    default: break;   // cylinder: no decode case in the reference (falls into its ASSERT); hull: decode returns false (:234-237)
                    }
                }
        }
    }
    return created;
}

} // namespace b200host
