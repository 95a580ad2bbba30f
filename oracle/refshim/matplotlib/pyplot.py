"""Import stub (test infrastructure only): /root/reference/plot_fleet.py calls plt.rc / plt.style.use at import."""


def rc(*a, **k):
    return None


class _Style:
    def use(self, *a, **k):
        return None


style = _Style()
