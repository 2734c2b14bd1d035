"""TEST INFRASTRUCTURE ONLY -- stand-in for the `opendr` package (absent from this image and from the reference tree).

renderer.py (the reference's mesh visualiser) imports opendr.camera.ProjectPoints, opendr.renderer.ColoredRenderer and
opendr.lighting.LambertianPointLight.  These modules give those three names the behaviour OpenDR documents, implemented
by the float64 functions of oracle/np_oracle.py (vert_normals, lambertian_point_light, rasterise), so that
oracle/make_render_vectors.py can run renderer.py itself, unmodified.  What that pins: the reference's own call sites --
camera defaults, near / far, the three lights, the albedo, the PLY colours of render_seg, the background image, the alpha
helpers and the uint8 conversion.  What it cannot pin: OpenDR's OpenGL rasterisation itself (see np_oracle.py).
"""
