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


# ---- objs/staircase: the material library and texture set of the reference's staircase scene -------------------------
#
# The reference ships objs/staircase/staircase.mtl (26 materials, Blender export) and ten JPEG textures (27.6 M texels, 83 MB as
# RGB8) but not the mesh.  The material table below restates the MTL's values in file order; the textures cannot be redistributed,
# so `write_staircase_assets` synthesises images with the reference files' names, dimensions and JPEG encodings (baseline or
# progressive, chroma subsampling, restart interval).  tests/test_host_scene.py checks the generated library against the
# reference's own file when /root/reference is present.
#            name                        Ns      Kd                   Ks                 Ni   d    illum  map_Kd
STAIRCASE_MATERIALS = [
    ("Black",                    250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Black.002",                250.0, (0.6, 0.6, 0.6),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Brass",                    250.0, (0.9, 0.8, 0.5),     (3.0, 2.0, 1.0),   1.0, 1.0, 4, None),
    ("Brushed_Aluminium",        250.0, (0.9, 0.7, 0.4),     (4.0, 3.0, 2.0),   1.0, 1.0, 4, "BrushedAluminium.jpg"),
    ("Brushed_Stainless_Steel",  250.0, (0.9, 0.7, 0.4),     (4.0, 3.0, 2.0),   1.0, 1.0, 4, None),
    ("Candles",                  250.0, (0.9, 0.8, 0.5),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Chair_Seat",                 0.0, (0.09, 0.02, 0.08),  (0.0, 0.0, 0.0),   1.0, 1.0, 2, "Fabric.jpg"),
    ("Emission",                 250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Glass",                    250.0, (1.0, 1.0, 1.0),     (1.0, 1.0, 1.0),   1.1, 0.1, 2, None),
    ("Gold",                     250.0, (0.2, 0.4, 1.4),     (3.4, 2.3, 1.8),   1.0, 1.0, 4, None),
    ("Lampshade_material",       250.0, (0.8, 0.65, 0.45),   (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Material.001",               0.0, (0.02, 0.05, 0.03),  (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Paint_-_Magnolia_Matt",    250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Paint_-_White_Gloss",      250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Paint_-_White_Matt",       250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Painting",                 250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, "Painting3.jpg"),
    ("Painting_01",              250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, "Painting1.jpg"),
    ("Painting_02",              250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, "Painting2.jpg"),
    ("Wallpaper",                250.0, (0.9, 0.8, 0.6),     (0.3, 0.3, 0.3),   1.0, 1.0, 2, "Wallpaper.jpg"),
    ("White.001",                250.0, (1.0, 1.0, 1.0),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("White_Plastic",            250.0, (0.8, 0.8, 0.8),     (0.0, 0.0, 0.0),   1.0, 1.0, 2, None),
    ("Wood_-_Chair",             250.0, (0.8, 0.3, 0.1),     (1.0, 1.0, 1.0),   1.0, 1.0, 2, "WoodChair.jpg"),
    ("Wood_-_Floor",            1000.0, (0.8, 0.3, 0.1),     (1.0, 1.0, 1.0),   1.0, 1.0, 2, "WoodFloor.jpg"),
    ("Wood_-_Lamp",              250.0, (0.6, 0.4, 0.1),     (1.0, 1.0, 1.0),   1.0, 1.0, 2, "Wood.jpg"),
    ("Wood_-_Stairs",           1000.0, (0.8, 0.3, 0.1),     (1.0, 1.0, 1.0),   1.0, 1.0, 2, "WoodPanel.jpg"),
    ("copper.001",               250.0, (0.9, 0.9, 0.1),     (0.1, 0.1, 0.1),   1.0, 1.0, 4, None),
]
# file name -> (width, height, chroma subsampling as PIL names it (0 = 4:4:4, 2 = 4:2:0), progressive, restart interval in MCUs)
STAIRCASE_TEXTURES = {
    "BrushedAluminium.jpg": (3500, 2625, 0, False, 438),
    "Fabric.jpg": (2048, 1382, 2, False, 0),
    "Painting1.jpg": (1024, 1024, 2, False, 0),
    "Painting2.jpg": (1440, 1440, 2, False, 0),
    "Painting3.jpg": (1920, 1920, 2, False, 0),
    "Wallpaper.jpg": (512, 512, 2, False, 0),
    "Wood.jpg": (1024, 689, 0, False, 128),
    "WoodChair.jpg": (3000, 2139, 0, False, 375),
    "WoodFloor.jpg": (960, 870, 2, False, 0),
    "WoodPanel.jpg": (726, 821, 0, True, 0),
}
STAIRCASE_FULL = os.path.join(ASSETS, "_generated", "staircase_full")   # git-ignored; written on first use


def staircase_mtl_text():
    lines = ["# objs/staircase/staircase.mtl: the reference's material library, restated from craytracer_b200/scenes.py", ""]
    for name, ns, kd, ks, ni, d, illum, tex in STAIRCASE_MATERIALS:
        lines += [f"newmtl {name}", f"Ns {ns}", "Ka 1.0 1.0 1.0", "Kd {} {} {}".format(*kd), "Ks {} {} {}".format(*ks), "Ke 0.0 0.0 0.0", f"Ni {ni}", f"d {d}"]
        if tex:
            lines.append(f"map_Kd textures/{tex}")
        lines += [f"illum {illum}", ""]
    return "\n".join(lines)


def write_staircase_assets(base_dir=STAIRCASE_FULL):
    """Writes <base_dir>/objs/staircase/staircase.mtl and its ten JPEG textures (synthetic content, the reference's dimensions and
    encodings) unless they are already there; returns base_dir, to be passed as the scene's base directory.  Needs PIL to encode."""
    import numpy as np
    from PIL import Image
    root = os.path.join(base_dir, "objs", "staircase")
    os.makedirs(os.path.join(root, "textures"), exist_ok=True)
    mtl = os.path.join(root, "staircase.mtl")
    if not os.path.exists(mtl) or open(mtl).read() != staircase_mtl_text():
        with open(mtl, "w") as f:
            f.write(staircase_mtl_text())
    for k, (name, (w, h, subsampling, progressive, restart)) in enumerate(sorted(STAIRCASE_TEXTURES.items())):
        path = os.path.join(root, "textures", name)
        if os.path.exists(path):
            continue
        ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
        rng = np.random.default_rng(1000 + k)
        f1, f2, ph = rng.uniform(0.004, 0.03, 2), rng.uniform(0.05, 0.2, 2), rng.uniform(0, 6.28, 3)
        grain = np.sin(xs * f1[0] + 6.0 * np.sin(ys * f1[1] + ph[0]) + ph[1]) * 0.5 + 0.5          # wood-grain-like bands
        weave = (np.sin(xs * f2[0]) * np.sin(ys * f2[1] + ph[2])) * 0.5 + 0.5                        # fine weave
        base = rng.uniform(0.25, 0.9, 3).astype(np.float32)
        img = np.stack([(0.55 * grain + 0.25 * weave + 0.2) * base[c] for c in range(3)], axis=-1)
        img = np.clip(img * 255.0 + rng.normal(0.0, 2.0, img.shape).astype(np.float32), 0, 255).astype(np.uint8)
        tmp = path + ".tmp"
        Image.fromarray(img).save(tmp, "JPEG", quality=88, subsampling=subsampling, progressive=progressive, optimize=progressive,
                                  **({"restart_marker_blocks": restart} if restart else {}))
        os.replace(tmp, path)
    return base_dir


def register_standins(dragon_triangles=DRAGON_TRIANGLES, interior_triangles=1_500_000):
    """Stand-ins for the meshes the reference does not ship; real files win when present under the base directory."""
    register_standin_mesh("objs/xyzrgb_dragon.obj", 0, dragon_triangles, 0)
    register_standin_mesh("objs/staircase/staircase.obj", 1, interior_triangles, 0)


CONFIGS = {"simple": simple, "cornell": cornell, "materials": materials, "dragon": dragon, "staircase": staircase, "anthropic": anthropic,
           "test": test_scene, "rounding-error": rounding_error}
