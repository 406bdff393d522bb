"""index.html over the rendered panels (reference utils/html.py:5-26)."""
import os


def make_html(home_dir):
    files = sorted(f for f in os.listdir(home_dir) if f.lower().endswith((".png", ".jpg")))
    with open(os.path.join(home_dir, "index.html"), "w") as f:
        f.write("<html><body><table>\n")
        for name in files:
            f.write('<tr><td>{0}<br><img src="{0}"></td></tr>\n'.format(name))
        f.write("</table></body></html>\n")
