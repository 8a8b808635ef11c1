import sys, importlib
sys.path.insert(0,'.')
scenes = importlib.import_module("path-tracing_b200.scenes"); core = importlib.import_module("path-tracing_b200.core")
for name in ("chess","dragon","street","atrium"):
    b,_,w,h,spp,d = scenes.WORKLOADS[name]
    with core.Renderer(0) as r:
        r.update_scene_data(b(w,h)); st=r.stats()
        print(name, "depth", st["bvh_max_depth"], "tris", st["triangle_count"], "refs", st["bvh_reference_count"], "nodes", st["bvh_node_count"], "GB %.2f"%(st["bvh_bytes"]/1e9), "build ms %.0f"%st["bvh_build_ms"])
