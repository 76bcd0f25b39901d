// OBJ / MTL loader -- same grammar and quirks as the reference loader (CG_Project/mesh.cpp:95-331
// loadMesh, :334-460 loadMtl), re-implemented from scratch around a small chunk reader.
//
// Grammar kept, with the reference line each rule comes from:
//   * input is consumed in chunks of at most 255 characters (fgets with LINE_LEN=256, mesh.cpp:22,151):
//     an over-long line is SPLIT and its tail is parsed as a line of its own.
//   * a chunk whose first character is '#', whitespace or NUL is skipped                   (mesh.cpp:154)
//   * "mtllib X": X is cut at the first character < 32 (signed char!) or == 255; the file is looked up
//     next to the OBJ; the directory string is APPENDED to, so a second mtllib resolves relative to the
//     first one's full path                                                                (mesh.cpp:157-178)
//   * "usemtl N": N ends at the first whitespace; unknown N -> warning, current material becomes ""
//                                                                                          (mesh.cpp:180-191)
//   * "v x y z" via sscanf("v %f %f %f") into variables that persist across lines          (mesh.cpp:193-197)
//   * "vt u v" stored, "vn" ignored                                                        (mesh.cpp:201-214)
//   * "f a b c ...": tokens are split on '/', ' ', CR, LF; the first component of every token is the
//     vertex id (atoi - 1, no negative/relative ids), the second the texcoord id; n-gons are fanned from
//     the first vertex (v0, v[i+1], v[i+2]); faces with < 3 vertices are dropped            (mesh.cpp:216-325)
//   * the triangle's material is the index of the current usemtl name; a name that is not in the table
//     (none yet, or unknown) is UB in the reference (find()==end() dereferenced, mesh.cpp:308,320) and is
//     PINNED to material 0 here.
//   * triangleMaterials is not cleared by loadMesh                                          (mesh.cpp:97-99)
//   * MTL: keys Kd Ka Ks Ns Ni illum map_Kd Tr d ("d" and "Tr" both write Tr); a material is committed on a
//     blank/indented line, on a failed read, or when EOF is hit while reading a statement; the first
//     definition of a name wins; newmtl itself does not commit the previous material; values leak from
//     one material to the next because cleanup() only clears flags                          (mesh.cpp:353-455)
#include "mesh.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

const int kChunk = 256;  // reference LINE_LEN (mesh.cpp:22): 255 payload characters + NUL

// Reads the next chunk the way fgets(buf, 256, f) would. Returns false when nothing could be read.
struct ChunkReader {
    FILE* f;
    char buf[kChunk];
    bool hit_eof;
    explicit ChunkReader(FILE* file) : f(file), hit_eof(false) { memset(buf, 0, sizeof(buf)); }
    bool next() {
        memset(buf, 0, sizeof(buf));
        int n = 0;
        while (n < kChunk - 1) {
            int c = fgetc(f);
            if (c == EOF) { hit_eof = true; break; }
            buf[n++] = (char)c;
            if (c == '\n') break;
        }
        return n > 0;
    }
};

inline bool is_space(char c) { return isspace((unsigned char)c) != 0; }
inline bool starts_with(const char* s, const char* key) { return strncmp(s, key, strlen(key)) == 0; }

// Text after the keyword (which is followed by at least one blank): leading whitespace skipped.
const char* after_keyword(const char* s, size_t keyword_len) {
    const char* p = s + keyword_len;
    while (is_space(*p)) ++p;
    return p;
}

// A name token: up to (not including) the first whitespace or the end of the chunk.
std::string name_token(const char* p) {
    const char* e = p;
    while (*e != '\0' && !is_space(*e)) ++e;
    return std::string(p, e);
}

// Splits the payload of an "f" statement into vertex-id and texcoord-id lists.
void parse_face(const char* payload, std::vector<int>& vids, std::vector<int>& tids) {
    vids.clear();
    tids.clear();
    const char* p = payload;
    while (*p == ' ') ++p;
    int component = 0;
    for (;;) {
        const char* tok = p;
        while (*p != '/' && *p != '\r' && *p != '\n' && *p != ' ' && *p != '\0') ++p;
        const bool slash = (*p == '/');
        if (p != tok) {
            std::string text(tok, p);
            if (component == 0) vids.push_back(atoi(text.c_str()) - 1);
            else if (component == 1) tids.push_back(atoi(text.c_str()) - 1);
            // component 2 (normal id) is ignored: normals are recomputed
        }
        component = slash ? component + 1 : 0;
        if (*p == '\0') break;
        ++p;  // step over the separator
        if (*p == '\0' || *p == '\n') break;
    }
}

}  // namespace

// Area-unweighted averaged vertex normals (reference mesh.cpp:28-47); preview-only, unused by the tracer.
void Mesh::computeVertexNormals() {
    for (size_t i = 0; i < vertices.size(); ++i) vertices[i].n = Vec3Df(0.f, 0.f, 0.f);
    for (size_t i = 0; i < triangles.size(); ++i) {
        const Vec3Df& a = vertices[triangles[i].v[0]].p;
        Vec3Df n = Vec3Df::crossProduct(vertices[triangles[i].v[1]].p - a, vertices[triangles[i].v[2]].p - a);
        n.normalize();
        for (int k = 0; k < 3; ++k) vertices[triangles[i].v[k]].n += n;
    }
    for (size_t i = 0; i < vertices.size(); ++i) vertices[i].n.normalize();
}

bool Mesh::loadMesh(const char* filename, bool /*randomizeTriangulation: the reference hard-wires k=0, mesh.cpp:297*/) {
    vertices.clear();
    triangles.clear();
    texcoords.clear();
    materials.clear();

    // Built-in material #0 (mesh.cpp:107-117): no Tr, no Ni.
    Material fallback;
    fallback.set_Kd(0.5f, 0.5f, 0.5f);
    fallback.set_Ka(0.f, 0.f, 0.f);
    fallback.set_Ks(0.5f, 0.5f, 0.5f);
    fallback.set_Ns(96.7f);
    fallback.set_illum(2);
    fallback.set_name("StandardMaterialInitFromTriMesh");
    materials.push_back(fallback);

    // Directory of the OBJ, with '\\' read as '/' (mesh.cpp:124-145).
    std::string dir(filename);
    for (size_t i = 0; i < dir.size(); ++i)
        if (dir[i] == '\\') dir[i] = '/';
    size_t slash = dir.rfind('/');
    dir = (slash == std::string::npos) ? std::string() : dir.substr(0, slash + 1);

    FILE* in = fopen(filename, "r");
    if (!in) return false;  // pinned: the reference calls fclose(NULL) here

    std::map<std::string, unsigned int> materialIndex;
    std::string current_material;  // "" until a known usemtl is seen
    std::vector<int> vids, tids;
    float x = 0.f, y = 0.f, z = 0.f;  // persist across "v" statements, as in the reference

    ChunkReader rd(in);
    while (!feof(in) && rd.next()) {
        char* s = rd.buf;
        if (s[0] == '#' || is_space(s[0]) || s[0] == '\0') continue;

        if (starts_with(s, "mtllib ")) {
            std::string rest = after_keyword(s, 6);
            size_t cut = 0;
            while (cut < rest.size() && !((signed char)rest[cut] < 32 || (unsigned char)rest[cut] == 255)) ++cut;
            dir.append(rest.substr(0, cut));  // yes: the directory string itself grows (mesh.cpp:173-175)
            if (verbose) printf("Load material file %s\n", dir.c_str());
            loadMtl(dir.c_str(), materialIndex);
        } else if (starts_with(s, "usemtl ")) {
            current_material = name_token(after_keyword(s, 6));
            if (materialIndex.find(current_material) == materialIndex.end()) {
                printf("Warning! Material '%s' not defined in material file. Taking default!\n", current_material.c_str());
                current_material.clear();
            }
        } else if (starts_with(s, "v ")) {
            sscanf(s, "v %f %f %f", &x, &y, &z);
            vertices.push_back(Vertex(Vec3Df(x, y, z)));
        } else if (starts_with(s, "vt ")) {
            Vec3Df tc(0.f, 0.f, 0.f);
            sscanf(s, "vt %f %f", &tc[0], &tc[1]);
            texcoords.push_back(tc);
        } else if (starts_with(s, "vn ")) {
            // face normals are recomputed by calculateNormals()
        } else if (starts_with(s, "f ")) {
            parse_face(s + 2, vids, tids);
            if (tids.size() != vids.size()) tids.resize(vids.size(), 0);
            std::map<std::string, unsigned int>::const_iterator it = materialIndex.find(current_material);
            const unsigned int m = (it == materialIndex.end()) ? 0u : it->second;  // pin (ii)
            const size_t n = vids.size();
            if (n >= 3) {
                for (size_t i = 0; i + 2 < n; ++i) {  // fan around the first vertex
                    triangles.push_back(Triangle(vids[0], tids[0], vids[i + 1], tids[i + 1], vids[i + 2], tids[i + 2]));
                    triangleMaterials.push_back(m);
                }
            } else {
                printf("TriMesh::LOAD: Unexpected number of face vertices (<3). Ignoring face");
            }
        }
    }
    fclose(in);
    return true;
}

bool Mesh::loadMtl(const char* filename, std::map<std::string, unsigned int>& materialIndex) {
    FILE* in = fopen(filename, "r");
    if (!in) {
        printf("  Warning! Material file '%s' not found!\n", filename);
        return false;
    }

    Material mat;              // one object for the whole file: cleanup() keeps its values (the leak)
    std::string key;           // name given by the last newmtl
    bool in_definition = false;  // set by the first newmtl and never cleared (mesh.cpp:350,384)
    float f1 = 0.f, f2 = 0.f, f3 = 0.f;  // persist across statements, as in the reference

    // Commits `mat` under `key` unless that name is already defined (first definition wins).
    struct Commit {
        static void run(std::vector<Material>& out, std::map<std::string, unsigned int>& index, Material& m, const std::string& k) {
            if (index.find(k) == index.end()) {
                m.set_name(k);
                out.push_back(m);
                index[k] = (unsigned int)out.size() - 1;
            }
        }
    };

    ChunkReader rd(in);
    while (!feof(in)) {
        rd.next();  // a failed read leaves an empty chunk, handled as "blank line at EOF" below
        const char* line = rd.buf;

        if (line[0] == '#') continue;

        if (is_space(line[0]) || line[0] == '\0') {
            if (in_definition && !key.empty() && mat.is_valid()) {
                Commit::run(materials, materialIndex, mat, key);
                mat.cleanup();
            }
            if (line[0] == '\0') break;
        } else if (starts_with(line, "newmtl ")) {
            key = name_token(after_keyword(line, 6));
            in_definition = true;
        } else if (starts_with(line, "Kd ")) {
            sscanf(line, "Kd %f %f %f", &f1, &f2, &f3);
            mat.set_Kd(f1, f2, f3);
        } else if (starts_with(line, "Ka ")) {
            sscanf(line, "Ka %f %f %f", &f1, &f2, &f3);
            mat.set_Ka(f1, f2, f3);
        } else if (starts_with(line, "Ks ")) {
            sscanf(line, "Ks %f %f %f", &f1, &f2, &f3);
            mat.set_Ks(f1, f2, f3);
        } else if (starts_with(line, "Ns ")) {
            sscanf(line, "Ns %f", &f1);
            mat.set_Ns(f1);
        } else if (starts_with(line, "Ni ")) {
            sscanf(line, "Ni %f", &f1);
            mat.set_Ni(f1);
        } else if (starts_with(line, "illum ")) {
            int illum = -1;
            sscanf(line, "illum %i", &illum);
            mat.set_illum(illum);
        } else if (starts_with(line, "map_Kd ")) {
            std::string t(line + 7);
            if (!t.empty() && t[t.size() - 1] == '\n') t.erase(t.size() - 1);
            mat.set_textureName(t);
        } else if (starts_with(line, "Tr ")) {
            sscanf(line, "Tr %f", &f1);
            mat.set_Tr(f1);
        } else if (starts_with(line, "d ")) {
            sscanf(line, "d %f", &f1);
            mat.set_Tr(f1);
        }

        // EOF reached while reading this statement (file without a trailing newline): commit now.
        if (feof(in) && in_definition && mat.is_valid() && !key.empty())
            Commit::run(materials, materialIndex, mat, key);
    }
    if (verbose) printf("%u  materials loaded.\n", (unsigned int)materials.size());
    fclose(in);
    return true;
}
