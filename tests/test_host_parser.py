"""Host logic: the product's .cry tokenizer / raw-value parser / typed conversion, restating the reference's
tests/test_parser.rs (same inputs, same expected tokens, error messages and line:column locations)."""
import os

import pytest

import craytracer_b200 as c


def kinds(text):
    return [(t["kind"], t.get("value")) for t in c.tokenize(text)]


def tok_error(text, message, location):
    with pytest.raises(c.ParserError) as e:
        c.tokenize(text)
    assert e.value.message == message
    assert e.value.location == location


def test_tokenizer_simple():  # test_parser.rs:31-64
    assert kinds("") == [("Eof", None)]
    assert kinds(" \r\t\n") == [("Eof", None)]
    assert kinds("// foo") == [("Eof", None)]
    assert kinds("{}") == [("LeftBrace", None), ("RightBrace", None), ("Eof", None)]
    assert kinds("[]") == [("LeftBracket", None), ("RightBracket", None), ("Eof", None)]
    assert kinds("()") == [("LeftParen", None), ("RightParen", None), ("Eof", None)]
    assert kinds("1") == [("Number", 1.0), ("Eof", None)]
    assert kinds("'hello'") == [("String", "hello"), ("Eof", None)]


def test_tokenizer_comments():  # :66-83
    assert kinds("// foo = 'hello'") == [("Eof", None)]
    assert kinds("//\n") == [("Eof", None)]
    assert kinds("//\n1") == [("Number", 1.0), ("Eof", None)]
    assert kinds("1 // one") == [("Number", 1.0), ("Eof", None)]
    tok_error("/", "Expected a second '/' to start a comment", (1, 2))
    tok_error("/ /", "Expected a second '/' to start a comment", (1, 2))


def test_tokenizer_numbers():  # :85-123
    for text, want in (("1", 1.0), ("1.0", 1.0), ("1.0000", 1.0), ("001.0000", 1.0), ("-1", -1.0), ("+1", 1.0), ("-1.1", -1.1), ("+1.1", 1.1)):
        assert kinds(text) == [("Number", want), ("Eof", None)]
    assert kinds("2+3") == [("Number", 2.0), ("Number", 3.0), ("Eof", None)]  # no expression support
    assert kinds("4-5") == [("Number", 4.0), ("Number", -5.0), ("Eof", None)]
    tok_error("9.8.7", "Unexpected character: '.'", (1, 4))


def test_tokenizer_strings():  # :125-147
    for s in ('""', "''", "'a'", '"Once upon a midnight dreary"', "'\"'", '"\'"'):
        assert kinds(s) == [("String", s[1:-1]), ("Eof", None)]


def test_tokenizer_identifiers():  # :149-162
    for s in ("simple", "snake_case", "camelCase", "SCREAMING_CASE", "agent007"):
        assert kinds(s) == [("Identifier", s), ("Eof", None)]


def test_tokenizer_locations():  # :164-235
    assert [(t["kind"], t["line"], t["column"]) for t in c.tokenize("{}")] == [("LeftBrace", 1, 1), ("RightBrace", 1, 2), ("Eof", 1, 3)]
    text = "\n{\n    x: 1,\n    y: ['foo', 3.14],\n}"
    got = [(t["kind"], t.get("value"), t["line"], t["column"]) for t in c.tokenize(text)]
    assert got == [("LeftBrace", None, 2, 1), ("Identifier", "x", 3, 5), ("Colon", None, 3, 6), ("Number", 1.0, 3, 8), ("Comma", None, 3, 9),
                   ("Identifier", "y", 4, 5), ("Colon", None, 4, 6), ("LeftBracket", None, 4, 8), ("String", "foo", 4, 9), ("Comma", None, 4, 14),
                   ("Number", 3.14, 4, 16), ("RightBracket", None, 4, 20), ("Comma", None, 4, 21), ("RightBrace", None, 5, 1), ("Eof", None, 5, 2)]


def test_tokenizer_full():  # :237-367 (token kinds of a whole scene)
    text = """
{
    camera: ProjectionCamera {
        origin: Point(0, 8, -10),
        up: Vector(0, 1, 0),
        fov: 5,
    },
    primitives: [
        Shape { shape: 'sky', material: 'sky' }
    ]
}
"""
    got = kinds(text)
    assert got[:16] == [("LeftBrace", None), ("Identifier", "camera"), ("Colon", None), ("Identifier", "ProjectionCamera"), ("LeftBrace", None),
                        ("Identifier", "origin"), ("Colon", None), ("Identifier", "Point"), ("LeftParen", None), ("Number", 0.0), ("Comma", None),
                        ("Number", 8.0), ("Comma", None), ("Number", -10.0), ("RightParen", None), ("Comma", None)]
    assert got[-1] == ("Eof", None) and got[-2] == ("RightBrace", None)
    assert ("String", "sky") in got


def raw_error(text, message, location):
    with pytest.raises(c.ParserError) as e:
        c.parse_raw_value(text)
    assert e.value.message == message
    assert e.value.location == location


def entries(v):
    return v["v"]["entries"]


def test_raw_value():  # :409-478
    assert c.parse_raw_value("1.23") == {"t": "Number", "v": 1.23}
    assert c.parse_raw_value("'hello'") == {"t": "String", "v": "hello"}
    assert c.parse_raw_value("Vector(1, -2, 3.1)") == {"t": "Vector", "v": [1.0, -2.0, 3.1]}
    assert c.parse_raw_value("Color(0, 0.5, 1)") == {"t": "Color", "v": [0.0, 0.5, 1.0]}
    v = c.parse_raw_value("{}")
    assert v["t"] == "Map" and entries(v) == {} and (v["v"]["line"], v["v"]["column"]) == (1, 1)
    v = c.parse_raw_value("{ x: 1, y: 'z' }")
    assert entries(v) == {"x": {"t": "Number", "v": 1.0}, "y": {"t": "String", "v": "z"}}
    v = c.parse_raw_value("Sphere { center: Point(0, 0, 0), radius: 1000 }")
    assert v["t"] == "TypedMap" and v["name"] == "Sphere" and (v["v"]["line"], v["v"]["column"]) == (1, 8)
    assert entries(v) == {"center": {"t": "Point", "v": [0.0, 0.0, 0.0]}, "radius": {"t": "Number", "v": 1000.0}}
    assert c.parse_raw_value("[]") == {"t": "Array", "v": []}
    v = c.parse_raw_value("[1, 'foo', {}]")
    assert v["v"][0] == {"t": "Number", "v": 1.0} and v["v"][1] == {"t": "String", "v": "foo"}
    assert (v["v"][2]["v"]["line"], v["v"][2]["v"]["column"]) == (1, 12)
    raw_error("x", "Expected '(' or '{', got EOF", (1, 2))
    raw_error(",", "Expected a raw value. Got ','", (1, 1))


def test_raw_value_map():  # :480-530
    assert entries(c.parse_raw_value("{ hello: 'world' }")) == {"hello": {"t": "String", "v": "world"}}
    assert entries(c.parse_raw_value("{ hello: 'world', }")) == {"hello": {"t": "String", "v": "world"}}
    e = entries(c.parse_raw_value("{ x: 1, y: 'z', v: Vector(1,2,3), c: Color(1,0,0) }"))
    assert e["v"] == {"t": "Vector", "v": [1.0, 2.0, 3.0]} and e["c"] == {"t": "Color", "v": [1.0, 0.0, 0.0]}
    raw_error("{ x: 1, x: 2 }", "Duplicate key x", (1, 1))
    raw_error("{ x: 1", "Expected '}', got EOF", (1, 7))
    raw_error("{ x 1 }", "Expected ':', got '1'", (1, 5))
    raw_error("{ 1: x }", "Expected '}', got '1'", (1, 3))
    raw_error("{ x: 1 y: 2 }", "Expected '}', got 'y'", (1, 8))


def test_raw_value_array():  # :532-584
    assert [x["v"] for x in c.parse_raw_value("[1, 2, 3]")["v"]] == [1.0, 2.0, 3.0]
    assert [x["v"] for x in c.parse_raw_value("[1, 2, 3,]")["v"]] == [1.0, 2.0, 3.0]
    raw_error("[", "Expected a raw value. Got EOF", (1, 2))
    raw_error("[,]", "Expected a raw value. Got ','", (1, 2))
    raw_error("[1 2]", "Expected ']', got '2'", (1, 4))


SCENE = """
{
    // Comment
    max_depth: 3,
    num_samples: 1,
    camera: Perspective {
        origin: Point(0, 0, 0),
        target: Point(0, 0, 1),
        up: Vector(0, 1, 0),
        fov: 60,
        lens_radius: 1,
        focal_distance: 100,
        film: {
            width: 400,
            height: 300
        },
    },
    lights: [
        Point {
            origin: Point(0, 0, 0),
            intensity: Color(1, 1, 1)
        }
    ],
    materials: {
        matte: Matte {
            reflectance: Color(1, 1, 1),
            sigma: 0
        },
        // Textures
        checks: Matte {
            reflectance: Checkerboard { a: Color(1, 1, 1), b: Color(0, 0, 0), scale: 2.5 },
            sigma: Checkerboard { a: 0, b: 1 }
        },
    },
    shapes: {
        ball: Sphere {
            origin: Point(0, 0, 2),
            radius: 1
        }
    },
    primitives: [
       Shape { shape: 'ball', material: 'matte' },
       Mesh { file_name: 'objs/triangle.obj', fallback_material: 'checks' },
    ]
}
"""


def test_parse_scene(tmp_path):  # :586-640, with the contents checked as well
    os.makedirs(tmp_path / "objs")
    (tmp_path / "objs" / "triangle.obj").write_text("# Simple triangle, used in tests\nv 1 0 0\nv 0 1 0\nv 0 0 1\n\nf 1 2 3")
    hs = c.parse_scene(SCENE, base_dir=tmp_path)
    d = hs.desc
    assert (d.max_depth, d.num_samples) == (3, 1)
    assert (d.camera.width, d.camera.height, d.camera.fov, d.camera.lens_radius, d.camera.focal_distance) == (400, 300, 60.0, 1.0, 100.0)
    assert d.n_lights == 1 and d.lights[0].kind == 0
    assert d.n_primitives == 2 and d.n_spheres == 1 and d.n_triangles == 1
    assert d.primitives[0].shape_kind == 0 and d.primitives[1].shape_kind == 1
    t = d.triangles[0]
    # RH -> LH flip (z -> -z) and v0, e1, e2 (src/obj.rs:126-133, src/shape.rs:108-109)
    assert list(t.v0) == [1.0, 0.0, -0.0] and list(t.e1) == [-1.0, 1.0, 0.0] and list(t.e2) == [-1.0, 0.0, -1.0]
    checks = d.materials[d.primitives[1].material]
    assert checks.t0.kind == 1 and checks.t0.scale == 2.5 and checks.t2.kind == 1 and checks.t2.scale == 1.0 and list(checks.t2.b)[0] == 1.0


def scene_error(text, message, location=None, base_dir="."):
    with pytest.raises(c.ParserError) as e:
        c.parse_scene(text, base_dir=base_dir)
    assert e.value.message == message, e.value.message
    if location is not None:
        assert e.value.location == location


MINIMAL = "{ camera: Perspective { origin: Point(0,0,0), target: Point(0,0,1), up: Vector(0,1,0), fov: 60, film: { width: 4, height: 4 } }, %s }"


def test_parse_scene_errors():  # error paths of scene_parser.rs:775-1117
    scene_error(MINIMAL % "lights: [], materials: {}, shapes: {}, primitives: []", "No lights in the scene.", (0, 0))
    scene_error("{ lights: [] }", "camera not found in map", (1, 1))
    scene_error(MINIMAL % "lights: [ Spot { } ], materials: {}, shapes: {}, primitives: []",
                "Error converting map value for 'lights' to expected type: Unknown light type: Spot")
    scene_error(MINIMAL % "lights: [], materials: { m: Velvet { } }, shapes: {}, primitives: []",
                "Error converting map value for 'materials' to expected type: Unknown material type: Velvet")
    scene_error(MINIMAL % "lights: [], materials: {}, shapes: {}, primitives: [ Shape { shape: 'nope', material: 'm' } ]",
                "Error converting map value for 'primitives' to expected type: Cannot find shape named 'nope'")
    scene_error(MINIMAL % "lights: [], materials: {}, shapes: { t: Triangle { v0: Point(0,0,0), v1: Point(1,0,0), v2: Point(2,0,0) } }, primitives: []",
                "Error converting map value for 'shapes' to expected type: Degenerate triangle: Triangle")
    scene_error(MINIMAL % "lights: [ Infinite { intensity: 3 } ], materials: {}, shapes: {}, primitives: []",
                "Error converting map value for 'lights' to expected type: Error converting map value for 'intensity' to expected type: Cannot get Color, found Number(3.0)")


def test_scene_defaults():  # DEFAULT_MAX_DEPTH 8, DEFAULT_NUM_SAMPLES 4, DEFAULT_FOCAL_DISTANCE 1e6 (scene_parser.rs:796-798)
    hs = c.parse_scene(MINIMAL % "lights: [ Infinite { intensity: Color(1,1,1) } ], materials: { m: Matte { reflectance: Color(1,1,1), sigma: 0 } }, "
                                 "shapes: { s: Sphere { origin: Point(0,0,3), radius: 1 } }, primitives: [ Shape { shape: 's', material: 'm', bogus: 1 } ]")
    d = hs.desc
    assert (d.max_depth, d.num_samples, d.camera.lens_radius, d.camera.focal_distance) == (8, 4, 0.0, 1e6)
    assert any("unused key" in w and "bogus" in w for w in hs.warnings)  # unused keys only warn (:571-586)


def test_area_lights_follow_explicit_lights_in_primitive_order():  # scene_parser.rs:1088-1101
    from craytracer_b200 import scenes
    hs = c.parse_scene(scenes.simple())  # the description's arrays live as long as the HostScene
    d = hs.desc
    assert d.n_lights == 2 and d.lights[0].kind == 2 and d.lights[1].kind == 3 and d.lights[1].primitive == 2
    assert d.primitives[2].area_light == 1 and d.primitives[0].area_light == -1
    hs2 = c.parse_scene(scenes.materials())
    d = hs2.desc
    assert d.n_primitives == 18 and d.lights[1].primitive == 0 and list(d.lights[1].color) == [8.0, 8.0, 8.0]
