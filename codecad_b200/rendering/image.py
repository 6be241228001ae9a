"""`render_pil_image` / `render_image` — /root/reference/codecad/rendering/image.py:7-20."""
from . import bitmap, ray_caster


def render_pixels(obj, size=(1024, 768), view_angle=None):
    """The uint8 [height][width][3] array behind render_pil_image (no PIL needed)."""
    if obj.dimension() == 2:
        return bitmap.render(obj, size)
    camera_params = ray_caster.get_camera_params(obj.bounding_box(), size, view_angle)
    return ray_caster.render(obj, size=size, *camera_params)


def render_pil_image(obj, size=(1024, 768), view_angle=None):
    import PIL.Image
    return PIL.Image.fromarray(render_pixels(obj, size, view_angle))


def render_image(obj, filename, size=(1024, 768), view_angle=None):
    render_pil_image(obj, size, view_angle).save(filename)
