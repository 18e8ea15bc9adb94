/*
 * flat_dump.h -- canonical byte dump of a FrayGpuScene (include/fray_gpu.h). TEST INFRASTRUCTURE.
 *
 * Used on both sides of the drop-in boundary: by oracle/ref_binding/ref_gpu_main.cpp (the reference's own Scene flattened by
 * FlatBuilder) and by oracle/ref_binding/host_flat_dump.cpp (this repository's host layer, fray_b200/host/flatten.cpp).
 * tests/test_ref_binding.py requires the two dumps to be equal byte for byte: the tables the GPU is handed do not depend on
 * which object model produced them. Every struct is memset to zero before it is filled on both sides, so padding compares too.
 */
#ifndef FRAY_FLAT_DUMP_H
#define FRAY_FLAT_DUMP_H
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include "fray_gpu.h"

static void frayDumpSection(FILE* f, const char* tag, const void* data, size_t bytes)
{
	char name[16];
	memset(name, 0, sizeof(name));
	strncpy(name, tag, sizeof(name) - 1);
	uint64_t n = (uint64_t) bytes;
	fwrite(name, 1, sizeof(name), f);
	fwrite(&n, sizeof(n), 1, f);
	if (bytes) fwrite(data, 1, bytes, f);
}

static int frayDumpFlat(const FrayGpuScene* s, const char* path)
{
	FILE* f = fopen(path, "wb");
	if (!f) return -1;
	frayDumpSection(f, "settings", &s->settings, sizeof(s->settings));
	frayDumpSection(f, "camera", &s->camera, sizeof(s->camera));
	frayDumpSection(f, "nodes", s->nodes, sizeof(FrayGpuNode) * (size_t) s->num_nodes);
	frayDumpSection(f, "geometries", s->geometries, sizeof(FrayGpuGeometry) * (size_t) s->num_geometries);
	frayDumpSection(f, "meshes", s->meshes, sizeof(FrayGpuMesh) * (size_t) s->num_meshes);
	frayDumpSection(f, "shaders", s->shaders, sizeof(FrayGpuShader) * (size_t) s->num_shaders);
	frayDumpSection(f, "layers", s->layers, sizeof(FrayGpuLayer) * (size_t) s->num_layers);
	frayDumpSection(f, "textures", s->textures, sizeof(FrayGpuTexture) * (size_t) s->num_textures);
	frayDumpSection(f, "bitmaps", s->bitmaps, sizeof(FrayGpuBitmap) * (size_t) s->num_bitmaps);
	frayDumpSection(f, "lights", s->lights, sizeof(FrayGpuLight) * (size_t) s->num_lights);
	frayDumpSection(f, "vertices", s->vertices, sizeof(double) * 3 * (size_t) s->num_vertices);
	frayDumpSection(f, "normals", s->normals, sizeof(double) * 3 * (size_t) s->num_normals);
	frayDumpSection(f, "uvs", s->uvs, sizeof(double) * 3 * (size_t) s->num_uvs);
	const size_t T = (size_t) s->num_triangles;
	frayDumpSection(f, "tri_v", s->tri_v, sizeof(int32_t) * 3 * T);
	frayDumpSection(f, "tri_n", s->tri_n, sizeof(int32_t) * 3 * T);
	frayDumpSection(f, "tri_t", s->tri_t, sizeof(int32_t) * 3 * T);
	frayDumpSection(f, "tri_gnormal", s->tri_gnormal, sizeof(double) * 3 * T);
	frayDumpSection(f, "tri_dndx", s->tri_dndx, sizeof(double) * 3 * T);
	frayDumpSection(f, "tri_dndy", s->tri_dndy, sizeof(double) * 3 * T);
	frayDumpSection(f, "tri_ab", s->tri_ab, sizeof(double) * 3 * T);
	frayDumpSection(f, "tri_ac", s->tri_ac, sizeof(double) * 3 * T);
	frayDumpSection(f, "tri_abxac", s->tri_abxac, sizeof(double) * 3 * T);
	frayDumpSection(f, "kd_nodes", s->kd_nodes, sizeof(FrayGpuKdNode) * (size_t) s->num_kd_nodes);
	frayDumpSection(f, "leaf_refs", s->leaf_refs, sizeof(int32_t) * (size_t) s->num_leaf_refs);
	frayDumpSection(f, "texels", s->texels, sizeof(float) * 3 * (size_t) s->num_texels);
	int32_t env[8];
	memset(env, 0, sizeof(env));
	env[0] = s->has_environment;
	for (int i = 0; i < 6; i++) env[1 + i] = s->env_bitmaps[i];
	frayDumpSection(f, "environment", env, sizeof(env));
	fclose(f);
	return 0;
}
#endif
