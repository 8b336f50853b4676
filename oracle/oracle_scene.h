/* oracle_scene.h -- data structures shared by the two oracle translation units
 * (ref_restate.c, path_oracle.c).
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, loaded or called
 * by the product (2019global_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
 */
#ifndef ORACLE_SCENE_H
#define ORACLE_SCENE_H

#include <stdint.h>
#include "g19.h"

typedef struct { double x, y, z; } d3;
typedef struct { float x, y, z; } f3;

/* ImpTriangle with the members the reference derives at construction
 * (reference include/entities.h:138-148, 251-254). */
typedef struct {
    d3 p1, p2, p3;
    d3 pos, edge1, edge2, normal;
    d3 bmin, bmax; /* ImpTriangle::min / ::max (max.z carries the +0.01) */
} o_tri;

typedef struct {
    int kind;
    d3 pos;        /* Entity::pos after construction */
    d3 color;      /* Material::color */
    float radius;  /* spheres */
    float f[4];    /* ctor scalars as given */
    d3 bbmin, bbmax; /* Entity::boundingBox() */
    int ntri;
    o_tri* tris;   /* composite triangles / own triangle(s) / box face pairs */
    d3 p3, p4;     /* ExpRectangle extras for getTextureCoord */
    int in_tree;   /* 0: rejected by Octree::push_back's root test (octree.h:22-24) */
    g19_entity_desc desc;
} o_entity;

typedef struct o_node {
    d3 bmin, bmax;
    int n_ent, cap_ent;
    int* ent;
    struct o_node* child[8];
} o_node;

typedef struct {
    d3 rmin, rmax;
    o_node* root;
    int n_ent, cap;
    o_entity* ent;
    int shapes; /* enum g19_shapes: which triangles the PATH oracle extracts from composites (REF restatement: unaffected) */
} o_scene;

#endif
