// host_flat_dump.cpp -- TEST INFRASTRUCTURE: dump the FrayGpuScene this repository's host layer (libfray_host.so:
// fray_b200/host/parser.cpp ... flatten.cpp) produces for a scene file, in the canonical form of flat_dump.h.
//   host_flat_dump scene.fray out.flat
#include <stdio.h>
#include "fray_host.h"
#include "flat_dump.h"

int main(int argc, char** argv)
{
	if (argc != 3) { fprintf(stderr, "usage: host_flat_dump scene.fray out.flat\n"); return 2; }
	FrayHostScene* scene = fray_host_load_scene(argv[1]);
	if (!scene) { fprintf(stderr, "%s\n", fray_host_last_error()); return 3; }
	const int rc = frayDumpFlat(fray_host_flat_scene(scene), argv[2]);
	fray_host_free_scene(scene);
	return rc ? 4 : 0;
}
