"""The benchmark scene descriptions of BASELINE.json, emitted as `.cry` text.

The parameters are those of the reference's scene files (scenes/*.cry); they are regenerated here from
tables instead of being shipped as copies.  Meshes that are not redistributable with the reference
(`.MISSING_LARGE_BLOBS`, git-ignored `objs/local`) are replaced by documented stand-ins:

* dragon    -- objs/xyzrgb_dragon.obj missing -> procedural closed tube, 7 219 045 triangles (register_standins)
* staircase -- objs/staircase/staircase.obj missing -> procedural interior using staircase.mtl next to it
* cornell   -- objs/local/cornell/CornellBox-Original.obj missing -> assets/cornell/CornellBox-Original.obj
               authored from the public Cornell box data (see tools/make_cornell.py)
"""
import os

from . import register_standin_mesh

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(REPO_ROOT, "assets")

DRAGON_TRIANGLES = 7_219_045


def _vec(kind, v):
    return f"{kind}({_num(v[0])}, {_num(v[1])}, {_num(v[2])})"


def _num(x):
    # the .cry number grammar has no exponent form
    if float(x) == int(x):
        return str(int(x))
    return format(float(x), "f").rstrip("0").rstrip(".") if abs(x) < 1e15 else str(x)


def _camera(origin, target, up, fov, width, height, lens_radius=None, focal_distance=None):
    extra = ""
    if lens_radius is not None:
        extra += f" lens_radius: {_num(lens_radius)},"
    if focal_distance is not None:
        extra += f" focal_distance: {_num(focal_distance)},"
    return (f"camera: Perspective {{ origin: {_vec('Point', origin)}, target: {_vec('Point', target)}, up: {_vec('Vector', up)}, "
            f"fov: {_num(fov)},{extra} film: {{ width: {width}, height: {height} }} }}")


def simple(num_samples=256, width=700, height=400):
    """scenes/simple.cry: ground disk, glass sphere, disk area light, infinite light."""
    return f"""{{
  num_samples: {num_samples},
  {_camera((-7.5, 6, -2), (-2.5, -1, 12), (0, 1, 0), 35, width, height)},
  lights: [ Infinite {{ intensity: Color(0.02, 0.08, 0.6) }} ],
  materials: {{
    ground: Matte {{ reflectance: Color(0.8, 0.8, 0.8), sigma: 0 }},
    glass: Glass {{ reflectance: Color(1, 1, 1), transmittance: Color(0.6, 0.6, 0.6), eta: 1.75 }}
  }},
  shapes: {{
    ground: Disk {{ origin: Point(0, 0, 10), radius: 40, rotate_x: 90 }},
    glass: Sphere {{ origin: Point(0, 1.5, 12.5), radius: 1.5 }},
    light: Disk {{ origin: Point(5, 5, 15), rotate_y: -40, rotate_x: 90, radius: 2 }}
  }},
  primitives: [
    Shape {{ shape: 'ground', material: 'ground' }},
    Shape {{ shape: 'glass', material: 'glass' }},
    Shape {{ shape: 'light', emittance: Color(10, 7, 1.2) }},
  ]
}}"""


_MATERIALS_TABLE = [
    # name, material text, sphere centre (x, z)
    ("matte1", "Matte { reflectance: Color(0.36, 0.86, 0.54), sigma: 0 }"),
    ("matte2", "Matte { reflectance: Color(0.57, 0.52, 0.85), sigma: 5 }"),
    ("matte3", "Matte { reflectance: Color(0.8, 0.2, 0.5), sigma: 30 }"),
    ("matte4", "Matte { reflectance: Color(0.6, 0.7, 0.8), sigma: 60 }"),
    ("brass", "Metal { eta: Color(0.44400, 0.52700, 1.09400), k: Color(3.69500, 2.76500, 1.82900) }"),
    ("chrome", "Metal { eta: Color(0.944, 0.776, 0.373), k: Color(4.0, 3.0, 2.0) }"),
    ("copper", "Metal { eta: Color(0.27105, 0.67693, 1.31640), k: Color(3.60920, 2.62480, 2.29210) }"),
    ("gold", "Metal { eta: Color(0.18299, 0.42108, 1.37340), k: Color(3.42420, 2.34590, 1.77040) }"),
    ("glass", "Glass { reflectance: Color(1, 1, 1), transmittance: Color(0.83, 0.85, 0.84), eta: 1.5 }"),
    ("water", "Glass { reflectance: Color(0.7, 0.9, 1), transmittance: Color(0.8, 0.8, 0.8), eta: 1.325 }"),
    ("diamond", "Glass { reflectance: Color(0.5, 0.5, 0.5), transmittance: Color(0.8, 0.8, 0.8), eta: 2.4 }"),
    ("emerald", "Glass { reflectance: Color(0.12, 0.3, 0.18), transmittance: Color(0.68, 0.8, 0.7), eta: 1.56 }"),
    ("plastic1", "Plastic { diffuse: Color(0.9, 0.9, 0.2), specular: Color(1, 1, 1), roughness: 0 }"),
    ("plastic2", "Plastic { diffuse: Color(0.2, 0.2, 1), specular: Color(0.7, 0.7, 1), roughness: 1 }"),
    ("plastic3", "Plastic { diffuse: Color(0.9, 0.1, 0.1), specular: Color(1, 1, 1), roughness: 10 }"),
    ("plastic4", "Plastic { diffuse: Color(0.4, 0.8, 0.9), specular: Color(1, 1, 1), roughness: 90 }"),
]


def materials(num_samples=1000, width=640, height=416):
    """scenes/materials.cry: a 4 x 4 grid of unit spheres (matte / metal / glass / plastic rows) on a huge ground sphere."""
    mats = ["ground: Matte { reflectance: Color(1, 1, 1), sigma: 30 }"] + [f"{n}: {m}" for n, m in _MATERIALS_TABLE]
    shapes = ["light: Sphere { origin: Point(10, 10, 10), radius: 5 }", "ground: Sphere { origin: Point(0, -100000, 10), radius: 100000 }"]
    prims = ["Shape { shape: 'light', emittance: Color(8, 8, 8) }", "Shape { shape: 'ground', material: 'ground' }"]
    for i, (name, _) in enumerate(_MATERIALS_TABLE):
        row, col = divmod(i, 4)
        shapes.append(f"{name}: Sphere {{ origin: Point({-3 + 2 * col}, 0.5, {-2 + 2 * row}), radius: 0.5 }}")
        prims.append(f"Shape {{ shape: '{name}', material: '{name}' }}")
    return f"""{{
  num_samples: {num_samples},
  {_camera((-15, 15, 50), (0, -2, 1), (-0.05, 1, 0), 12.5, width, height)},
  lights: [ Infinite {{ intensity: Color(0.5, 0.5, 0.5) }} ],
  materials: {{ {", ".join(mats)} }},
  shapes: {{ {", ".join(shapes)} }},
  primitives: [ {", ".join(prims)} ]
}}"""


def dragon(num_samples=10, width=600, height=400, mesh="objs/xyzrgb_dragon.obj"):
    """scenes/dragon.cry: gold dragon on a huge ground sphere under a disk light."""
    return f"""{{
  num_samples: {num_samples},
  {_camera((150, 70, 150), (30, -50, 0), (0, 1, 0), 60, width, height)},
  lights: [],
  materials: {{
    ground: Matte {{ reflectance: Color(1, 1, 1), sigma: 0 }},
    dragon: Metal {{ eta: Color(0.18299, 0.42108, 1.37340), k: Color(3.42420, 2.34590, 1.77040) }}
  }},
  shapes: {{
    light: Disk {{ origin: Point(0, 80, 0), rotate_x: 90, radius: 50 }},
    ground: Sphere {{ origin: Point(0, -100040, 10), radius: 100000 }}
  }},
  primitives: [
    Shape {{ shape: 'ground', material: 'ground' }},
    Shape {{ shape: 'light', emittance: Color(1, 1, 1) }},
    Mesh {{ file_name: '{mesh}', fallback_material: 'dragon' }}
  ]
}}"""


def cornell(num_samples=16, width=400, height=400, mesh="objs/local/cornell/CornellBox-Original.obj"):
    """scenes/cornell.cry: the Cornell box mesh, lit only by its emissive quad."""
    return f"""{{
  num_samples: {num_samples},
  max_depth: 8,
  {_camera((0, 1, -2.8), (0, 1, 0), (0, 1, 0), 60, width, height)},
  lights: [],
  materials: {{ default: Matte {{ reflectance: Color(1, 1, 1), sigma: 0 }} }},
  shapes: {{}},
  primitives: [ Mesh {{ file_name: '{mesh}', fallback_material: 'default' }} ]
}}"""


def staircase(num_samples=64, width=720, height=1280, mesh="objs/staircase/staircase.obj"):
    """scenes/staircase.cry: textured interior, thin-lens camera, point light + disk area light."""
    return f"""{{
  num_samples: {num_samples},
  {_camera((0, 2, -4.92), (0, 2.5, 0), (0, 1, 0), 35, width, height, lens_radius=0.001, focal_distance=3)},
  lights: [ Point {{ origin: Point(0, 2.25, -4.5), intensity: Color(0.3, 0.3, 0.3) }} ],
  materials: {{ default: Matte {{ reflectance: Color(1, 1, 1), sigma: 0 }} }},
  shapes: {{ light1: Disk {{ origin: Point(1, 5.5, 2.5), rotate_x: 60, rotate_y: 0, radius: 2 }} }},
  primitives: [
    Shape {{ shape: 'light1', emittance: Color(5, 5, 5) }},
    Mesh {{ file_name: '{mesh}', fallback_material: 'default' }},
  ]
}}"""


def anthropic(num_samples=1024, width=800, height=600, mesh="objs/anthropic.obj"):
    """scenes/anthropic.cry: the only in-tree mesh scene of the reference (20 060 triangles)."""
    return f"""{{
  num_samples: {num_samples},
  {_camera((0.7, 2, 1.0), (0.3, 0, 0.05), (0, 0, 1), 30, width, height)},
  lights: [ Infinite {{ intensity: Color(1.6, 1.6, 1.5) }} ],
  materials: {{
    ground: Matte {{ reflectance: Color(1, 1, 1), sigma: 10 }},
    text: Plastic {{ diffuse: Color(0.03, 0.03, 0.03), specular: Color(0.2, 0.2, 0.2), roughness: 120 }}
  }},
  shapes: {{
    ground: Disk {{ origin: Point(0.25, 0, -0.001), radius: 20, rotate_x: 0 }},
    light: Disk {{ origin: Point(0.5, 0.1, 1), rotate_y: 0, rotate_x: 0, radius: 1 }}
  }},
  primitives: [
    Shape {{ shape: 'ground', material: 'ground' }},
    Mesh {{ file_name: '{mesh}', fallback_material: 'text' }},
    Shape {{ shape: 'light', emittance: Color(1, 0, 0) }},
  ]
}}"""


def test_scene(num_samples=256, width=500, height=500):
    """scenes/test.cry ("Test scene for judging correctness"): distant light, triangle light, checkerboard, all shape kinds."""
    return f"""{{
  num_samples: {num_samples},
  {_camera((1.5, 1.5, -4), (1.5, 1, 0), (0, 1, 0), 60, width, height)},
  lights: [ Distant {{ direction: Vector(0, 0, -1), intensity: Color(1, 1, 1) }} ],
  materials: {{
    white: Matte {{ reflectance: Color(1, 1, 1), sigma: 100 }},
    red: Matte {{ reflectance: Color(1, 0, 0), sigma: 0 }},
    green: Matte {{ reflectance: Color(0, 1, 0), sigma: 0 }},
    blue: Matte {{ reflectance: Checkerboard {{ a: Color(0, 0, 1), b: Color(1, 1, 1), scale: 4 }}, sigma: 0 }},
    mirror: Metal {{ eta: Color(0.9, 0.8, 0.4), k: Color(4.0, 3.0, 2.0) }},
    glass: Glass {{ reflectance: Color(1, 1, 1), transmittance: Color(0.9, 0.9, 0.9), eta: 1.5 }},
  }},
  shapes: {{
    light: Triangle {{ v0: Point(0, 0, 0), v1: Point(0, 0, -1), v2: Point(0, 1, 0) }},
    ground: Sphere {{ origin: Point(0, -100, 0), radius: 100 }},
    triangle1: Triangle {{ v0: Point(0, 0, 0), v1: Point(1, 0, 0), v2: Point(0, 1, 0) }},
    triangle2: Triangle {{ v0: Point(2, 0, 0), v1: Point(2, 1, 0), v2: Point(3, 0, 0) }},
    disk1: Disk {{ origin: Point(0.5, 2, 0), radius: 0.5, rotate_x: 180 }},
    disk2: Disk {{ origin: Point(2.5, 2, 0), radius: 0.5, rotate_x: 0 }},
    glass: Sphere {{ origin: Point(0.5, 0.25, -1), radius: 0.25 }},
    mirror: Sphere {{ origin: Point(2.5, 0.25, -1), radius: 0.25 }},
    sphere: Sphere {{ origin: Point(1.5, 0.25, -1), radius: 0.25 }},
  }},
  primitives: [
    Shape {{ shape: 'light', emittance: Color(1, 1, 1) }},
    Shape {{ shape: 'ground', material: 'white' }},
    Shape {{ shape: 'triangle1', material: 'red' }},
    Shape {{ shape: 'triangle2', material: 'green' }},
    Shape {{ shape: 'disk1', material: 'red' }},
    Shape {{ shape: 'disk2', material: 'green' }},
    Shape {{ shape: 'glass', material: 'glass' }},
    Shape {{ shape: 'mirror', material: 'mirror' }},
    Shape {{ shape: 'sphere', material: 'blue' }},
  ]
}}"""


def rounding_error(num_samples=10, width=400, height=400):
    """scenes/rounding-error.cry: documents the reference's shadow-ray / AABB false miss."""
    return f"""{{
  num_samples: {num_samples},
  max_depth: 1,
  {_camera((0, 5, -5), (0, -1, 5), (0, 1, 0), 60, width, height)},
  lights: [],
  materials: {{ ground: Matte {{ reflectance: Color(0.8, 0.8, 0.8), sigma: 0 }} }},
  shapes: {{
    ground: Disk {{ origin: Point(0, 0, 0), rotate_x: -90, radius: 10 }},
    ball: Sphere {{ origin: Point(0, 1.5, 2.5), radius: 1.5 }},
    light: Sphere {{ origin: Point(0, 5, 5), radius: 1 }}
  }},
  primitives: [
    Shape {{ shape: 'ground', material: 'ground' }},
    Shape {{ shape: 'ball', material: 'ground' }},
    Shape {{ shape: 'light', emittance: Color(5, 5, 5) }},
  ]
}}"""


def register_standins(dragon_triangles=DRAGON_TRIANGLES, interior_triangles=1_500_000):
    """Stand-ins for the meshes the reference does not ship; real files win when present under the base directory."""
    register_standin_mesh("objs/xyzrgb_dragon.obj", 0, dragon_triangles, 0)
    register_standin_mesh("objs/staircase/staircase.obj", 1, interior_triangles, 0)


CONFIGS = {"simple": simple, "cornell": cornell, "materials": materials, "dragon": dragon, "staircase": staircase, "anthropic": anthropic,
           "test": test_scene, "rounding-error": rounding_error}
