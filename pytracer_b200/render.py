"""The renderers of the reference (render.py:26-193) backed by the CUDA library.

``CudaRenderer`` is the plug-in: it is constructed like the reference's renderers
(``Cls(world=..., background_color=..., ...)``), it is a callable ``Ray -> Color`` like them — so it
can be handed to the reference's own ``ImageTracer.fire_all_rays`` and will work ray by ray — and
:class:`pytracer_b200.imagetracer.CudaImageTracer` recognises it and renders the whole image with
one ``rt_render`` call instead.  ``OnOffRenderer`` / ``FlatRenderer`` / ``PathTracer`` /
``PointLightRenderer`` keep the reference's names and signatures.

``world`` may be a :class:`pytracer_b200.scene.World` or the reference's ``pytracer.world.World``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _abi
from .device import DeviceScene, MultiDeviceScene
from .params import make_params
from .pcg import PCG
from .scene import BLACK, WHITE, Color, Ray

RENDERERS = ["onoff", "flat", "pathtracing", "pointlight"]  # main.py:73


class Renderer:
    """render.py:26-39"""

    def __init__(self, world, background_color: Color = BLACK):
        self.world = world
        self.background_color = background_color

    def __call__(self, ray: Ray) -> Color:
        raise NotImplementedError("Unable to call Renderer.radiance, it is an abstract method")


class CudaRenderer(Renderer):
    def __init__(self, world, background_color: Color = BLACK, algorithm: str = "pathtracing",
                 pcg: Optional[PCG] = None, num_of_rays: int = 10, max_depth: int = 10,
                 russian_roulette_limit: int = 3, ambient_color: Color = None, color: Color = WHITE,
                 variant: str = "auto", precision: str = "auto", accel: str = "none", gpus=1):
        """``gpus``: how many devices of this node render an image (or a list of device indices), all from
        this process and one ``rt_render_multi`` call per image — interleaved rows, the same image bit for bit.
        (One process per GPU under ``torchrun`` is the other way: ``fire_all_rays(..., comm=)``.)"""
        super().__init__(world, background_color)
        if algorithm not in RENDERERS:
            raise ValueError(f"Unknown renderer: {algorithm}")
        self.algorithm = algorithm
        self.pcg = pcg if pcg is not None else PCG()
        self.num_of_rays = num_of_rays
        self.max_depth = max_depth
        self.russian_roulette_limit = russian_roulette_limit
        self.ambient_color = ambient_color if ambient_color is not None else Color(0.1, 0.1, 0.1)
        self.color = color
        self.variant = variant
        self.precision = precision
        self.accel = accel  # "none": the reference's loop over all shapes; "bvh": sphere hierarchy (same image)
        self.gpus = gpus
        self.last_stats: dict = {}
        self._scene: Optional[DeviceScene] = None

    # -- device scene: the reference's renderers read the World live at every call, so the world is
    # flattened again for every image and compared with what is resident in HBM (bulk array compares)
    def device_scene(self) -> DeviceScene:
        """The resident device scene, brought up to date with ``self.world``: unchanged -> reused as is;
        only transformations edited -> patched in place (rt_scene_update_transforms); anything else
        (shapes added / replaced, materials, pigments, textures, lights edited) -> rebuilt."""
        from .flatten import flatten_world

        new = flatten_world(self.world)
        if self._scene is None:
            self._scene = self._new_scene(new)
        elif not self._scene.flat.differs_only_in_transforms(new):
            self._scene.close()
            self._scene = self._new_scene(new)
        elif not (np.array_equal(self._scene.flat.shape_m, new.shape_m) and np.array_equal(self._scene.flat.shape_invm, new.shape_invm)):
            self._scene.update_from_world(new)
        return self._scene

    def _new_scene(self, flat):
        multi = not isinstance(self.gpus, int) or self.gpus > 1
        return MultiDeviceScene(flat, self.gpus) if multi else DeviceScene(flat)

    def refresh(self) -> None:
        """Flatten and upload the world again unconditionally."""
        from .flatten import flatten_world

        if self._scene is not None:
            self._scene.close()
        self._scene = self._new_scene(flatten_world(self.world))

    def set_world(self, world) -> bool:
        """Next frame of an animation (the reference re-parses the scene with another `clock`,
        main.py:122-128): if only transformations changed the resident device scene is patched in
        place (rt_scene_update_transforms), otherwise it is rebuilt.  Returns True when patched."""
        had = self._scene
        self.world = world
        return self.device_scene() is had

    def make_params(self, width: int, height: int, camera, samples_per_side: int = 0, aa_pcg: Optional[PCG] = None,
                    **overrides) -> _abi.rt_render_params:
        kw = dict(
            algorithm=self.algorithm, samples_per_side=samples_per_side, background=self.background_color,
            onoff_color=self.color, ambient=self.ambient_color, num_of_rays=self.num_of_rays,
            max_depth=self.max_depth, rr_limit=self.russian_roulette_limit, aa_pcg=aa_pcg, pt_pcg=self.pcg,
            variant=self.variant, precision=self.precision, accel=self.accel,
        )
        kw.update(overrides)
        return make_params(width, height, camera, **kw)

    def __call__(self, ray: Ray) -> Color:
        """Renderer.__call__(ray): one ray through the device code, draws taken from ``self.pcg`` in
        the reference's order (the generator is left where the reference would leave it)."""
        from .scene import PerspectiveCamera

        params = self.make_params(1, 1, PerspectiveCamera(), precision="f64" if self.precision == "auto" else self.precision)
        rays = np.array([[ray.origin.x, ray.origin.y, ray.origin.z, ray.dir.x, ray.dir.y, ray.dir.z, ray.tmin, ray.tmax]])
        rgb, (state, _) = self.device_scene().trace_rays(params, rays, np.array([ray.depth], dtype=np.int32),
                                                         (self.pcg.state, self.pcg.inc))
        self.pcg.state = state
        return Color(*rgb[0])


class OnOffRenderer(CudaRenderer):
    """render.py:42-53"""

    def __init__(self, world, background_color: Color = BLACK, color: Color = WHITE, **kw):
        super().__init__(world, background_color, algorithm="onoff", color=color, **kw)


class FlatRenderer(CudaRenderer):
    """render.py:56-74"""

    def __init__(self, world, background_color: Color = BLACK, **kw):
        super().__init__(world, background_color, algorithm="flat", **kw)


class PathTracer(CudaRenderer):
    """render.py:77-139"""

    def __init__(self, world, background_color: Color = BLACK, pcg: Optional[PCG] = None, num_of_rays: int = 10,
                 max_depth: int = 10, russian_roulette_limit: int = 3, **kw):
        super().__init__(world, background_color, algorithm="pathtracing", pcg=pcg, num_of_rays=num_of_rays,
                         max_depth=max_depth, russian_roulette_limit=russian_roulette_limit, **kw)


class PointLightRenderer(CudaRenderer):
    """render.py:142-193"""

    def __init__(self, world, background_color: Color = BLACK, ambient_color: Color = None, **kw):
        super().__init__(world, background_color, algorithm="pointlight", ambient_color=ambient_color, **kw)
