"""Scene XML v1.1 -> (emitter_configs, array_info, all_objs, configs).

Same schema, same 4-tuple and same error behaviour as the reference ``parsers/xml_parser.py``
(scene_parsing :246-289, parse_wavefront :93-176, parse_global_sensor :225-244, parse_emitters
:66-88, parse_bxdf :178-194, update_emitter_config :56-64), re-hosted without taichi / pywavefront.
Textures (albedo / normal / bump atlases, parse_texture :203-221) are parsed like the reference does; volumes are outside
the `pt` hot path.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as xet
from typing import List

import numpy as np

from ..bxdf import brdf as _brdf_mod
from ..bxdf.brdf import BRDF_np
from ..bxdf.bsdf import BSDF_np
from ..bxdf.texture import Texture_np
from ..emitters.area import AreaSource
from ..emitters.collimated import CollimatedSource
from ..emitters.point import PointSource
from ..emitters.spot import SpotSource
from ..utils.tools import CONSOLE, timing
from .general_parser import get, parse_sphere_element, transform_parse
from .obj_desc import ObjDescriptor
from .texture_packing import image_packer
from .obj_loader import SPHERE, TRIANGLE_MESH, apply_transform, calculate_surface_area, extract_obj_info
from .world import World_np

__all__ = ["scene_parsing"]

__VERSION__ = "1.1"
__MAPPING__ = {"integer": int, "float": float, "string": str,
               "boolean": lambda x: True if x.lower() == "true" else False}
__SOURCE_MAP__ = {"point": PointSource, "area": AreaSource, "spot": SpotSource, "collimated": CollimatedSource}


def none_checker(value, prim_num, last_dim=3):
    if value is None:
        return np.zeros((prim_num, 3, last_dim), dtype=np.float32)
    return value


def update_emitter_config(emitter_config: List, area_lut: dict):
    for i, emitter in enumerate(emitter_config):
        if i in area_lut:
            emitter.inv_area = 1.0 / area_lut[i]
            emitter.attached = True
        elif emitter.type == "area":
            raise ValueError("Setting L1 / L2 for area light is deprecated a long ago. Please attach area light to an object.")
    return emitter_config


def parse_emitters(em_elem: list):
    sources = []
    source_id_dict = dict()
    for elem in em_elem:
        emitter_type = elem.get("type")
        source_type = __SOURCE_MAP__.get(emitter_type, None)
        if source_type is None:
            raise ValueError(f"Source type '{emitter_type}' is not supported. Please check your XML settings.")
        source = source_type(elem)
        if source.id in source_id_dict:
            raise ValueError(f"Two sources with same id {source.id} will result in conflicts")
        source_id_dict[source.id] = len(sources)
        sources.append(source)
    return sources, source_id_dict


def parse_wavefront(directory: str, obj_list: List[xet.Element], bsdf_dict: dict, emitter_dict: dict, texture_dict: dict = None):
    all_objs, all_prims, all_uvs, all_normals, all_v_norms = [], [], [], [], []
    indices = []
    attached_area_dict = {}
    has_vertex_normal = False
    cum_prim_num = 0
    for elem in obj_list:
        vns, uvs, trans_r, trans_t = None, None, None, None
        obj_type = TRIANGLE_MESH
        if elem.get("type") == "obj":
            filepath_child = elem.find("string")
            meshes, normals, vns, uvs = extract_obj_info(os.path.join(directory, filepath_child.get("value")))
            transform_child = elem.find("transform")
            if transform_child is not None:
                trans_r, trans_t, trans_s = transform_parse(transform_child)
                meshes, normals = apply_transform(meshes, normals, trans_r, trans_t, trans_s)
            if vns is not None:
                has_vertex_normal = True
        else:
            meshes, normals = parse_sphere_element(elem)
            obj_type = SPHERE
        bsdf_item = None
        texture_group = {"albedo": None, "normal": None, "bump": None, "roughness": None}
        emit_ref_id = -1
        for ref_child in elem.findall("ref"):
            ref_type = ref_child.get("type")
            ref_id = ref_child.get("id")
            if ref_type == "material":
                bsdf_item = bsdf_dict[ref_id]
            elif ref_type == "emitter":
                emit_ref_id = emitter_dict[ref_id]
                attached_area_dict[emit_ref_id] = calculate_surface_area(meshes, obj_type)
            elif ref_type == "texture":
                ref_tag = ref_child.get("tag", None)
                if ref_tag is None:
                    ref_tag = "albedo"
                    CONSOLE.log(f"[yellow]Warning: BXDF[/yellow] Texture ref_id {ref_id} has no tag. Set default as 'albedo'.")
                elif ref_tag not in texture_group:
                    ref_tag = "albedo"
                    CONSOLE.log(f"[yellow]Warning: BXDF[/yellow] Texture ref_tag {ref_tag} not supported. Set default as 'albedo'.")
                if texture_dict is None or texture_dict.get(ref_tag) is None or ref_id not in texture_dict[ref_tag]:
                    raise KeyError(f"Texture id '{ref_id}' does not have tag '{ref_tag}' mapping, check if it is from other groups.")
                texture_group[ref_tag] = texture_dict[ref_tag][ref_id]
                if texture_group[ref_tag].mode == Texture_np.MODE_CHECKER:
                    raise NotImplementedError("checkerboard textures have no lookup in the reference (bxdf/texture.py:102 TODO)")
        if bsdf_item is None:
            raise ValueError("Object should be attached with a BSDF for now since no default one implemented yet.")
        prim_num = meshes.shape[0]
        if obj_type == SPHERE:
            meshes = np.concatenate((meshes, np.zeros((1, 1, 3), dtype=np.float32)), axis=-2)
            indices.append(cum_prim_num)
        all_prims.append(meshes)
        all_normals.append(normals)
        all_v_norms.append(none_checker(vns, prim_num))
        all_uvs.append(none_checker(uvs, prim_num, last_dim=2))
        all_objs.append(ObjDescriptor(meshes, normals, bsdf_item, vns, uvs, texture_group, trans_r, trans_t,
                                      emit_ref_id, obj_type))
        cum_prim_num += prim_num
    indices = np.int64(indices) if indices else None
    array_info = {
        "primitives": np.concatenate(all_prims, axis=0).astype(np.float32),
        "indices": indices,
        "n_g": np.concatenate(all_normals, axis=0).astype(np.float32),
        "n_s": np.concatenate(all_v_norms, axis=0).astype(np.float32),
        "uvs": np.concatenate(all_uvs, axis=0).astype(np.float32),
    }
    return array_info, all_objs, attached_area_dict, has_vertex_normal


def parse_bxdf(bxdf_list: List[xet.Element]):
    results = dict()
    for bxdf_node in bxdf_list:
        bxdf_id = bxdf_node.get("id")
        bxdf = BRDF_np(bxdf_node) if bxdf_node.tag == "brdf" else BSDF_np(bxdf_node)
        if bxdf_id in results:
            CONSOLE.log(f"[yellow]Warning: BXDF[/yellow] {bxdf_id} re-defined in XML file. Overwriting the existing BXDF.")
        results[bxdf_id] = bxdf
    return results


def parse_texture(texture_list: List[xet.Element], directory: str = ""):
    """Texture nodes -> ({tag: atlas image or None}, {tag: {id: Texture_np} or None}) (reference xml_parser.py:203-221)."""
    if len(texture_list) == 0:
        return None, None
    textures = {"albedo": [], "normal": [], "bump": [], "roughness": []}
    for texture in texture_list:
        textures[texture.get("tag", "albedo")].append(Texture_np(texture, directory=directory))
    packed_textures, packed_imgs = {}, {}
    for key, value in textures.items():
        if len(value) == 0:
            tex_img, tex_info = None, None
        else:
            tex_img, tex_info = image_packer(value)
        packed_imgs[key] = tex_img
        packed_textures[key] = tex_info
    return packed_imgs, packed_textures


def parse_world(world_elem: xet.Element):
    world = World_np(world_elem)
    if world_elem is None:
        CONSOLE.log("[yellow]Warning: world element not found in xml file. Using default world settings:")
    return world


def parse_global_sensor(sensor_elem: xet.Element):
    sensor_config = {}
    for elem in sensor_elem:
        if elem.tag in __MAPPING__:
            sensor_config[elem.get("name")] = get(elem, "value", __MAPPING__[elem.tag])
    sensor_config["transform"] = transform_parse(sensor_elem.find("transform"))
    film_elems = sensor_elem.find("film").findall("integer")
    assert len(film_elems) >= 2
    sensor_config["film"] = {}
    for elem in film_elems:
        sensor_config["film"][elem.get("name")] = get(elem, "value", __MAPPING__[elem.tag])
    return sensor_config


@timing()
def scene_parsing(directory: str, file: str):
    xml_file = os.path.join(directory, file)
    CONSOLE.log(f":fax: Parsing XML file from '{xml_file}'")
    root_node = xet.parse(xml_file).getroot()
    version_tag = root_node.attrib["version"]
    if not version_tag == __VERSION__:
        raise ValueError(f"Unsupported version {version_tag}. Only '{__VERSION__}' is supported right now.")
    bxdf_nodes = root_node.findall("bsdf") + root_node.findall("brdf")
    texture_nodes = root_node.findall("texture")
    emitter_nodes = root_node.findall("emitter")
    shape_nodes = root_node.findall("shape")
    sensor_node = root_node.find("sensor")
    world_node = root_node.find("world")
    volume_node = root_node.findall("volume")
    assert sensor_node is not None
    teximgs, textures = parse_texture(texture_nodes, directory)
    # the reference flips microfacet support with a source-level flag (bxdf/brdf.py:8); here a sensor key, default off like the
    # reference's, and reset for every scene so that one scene's setting never leaks into the next one parsed by this process
    _brdf_mod.set_enable_microfacet(False)
    for elem in sensor_node:
        if elem.tag == "boolean" and elem.get("name") == "enable_microfacet":
            _brdf_mod.set_enable_microfacet(elem.get("value", "false").lower() == "true")
    emitter_configs, emitter_dict = parse_emitters(emitter_nodes)
    bsdf_dict = parse_bxdf(bxdf_nodes)
    array_info, all_objs, area_lut, has_vertex_normal = parse_wavefront(directory, shape_nodes, bsdf_dict, emitter_dict, textures)
    configs = parse_global_sensor(sensor_node)
    configs["world"] = parse_world(world_node)
    configs["packed_textures"] = teximgs
    configs["has_vertex_normal"] = has_vertex_normal
    configs["volume"] = volume_node[:1]
    emitter_configs = update_emitter_config(emitter_configs, area_lut)
    return emitter_configs, array_info, all_objs, configs
