import sys; sys.path.insert(0,'/root/repo')
from raytracingoneweekendapplication_b200 import capi
sc=capi.Scene("final"); c=capi.Context(0); c.upload(sc)
def t(w,h,spp,depth=50,**kw):
    b=1e9
    for _ in range(5):
        c.render(w,h,spp,max_depth=depth,seed=1,**kw); b=min(b,c.stats()["render_ms"])
    return round(b,4)
print("64x64 1spp", t(64,64,1))
print("512x512 1spp", t(512,512,1))
print("4K 1spp depth50", t(3840,2160,1))
print("4K 1spp depth1", t(3840,2160,1,depth=1))
print("4K 2spp", t(3840,2160,2), "4K 4spp", t(3840,2160,4), "4K 8spp", t(3840,2160,8))
print("quads 64x64", end=" ")
c.upload(capi.Scene("quads")); print(t(64,64,1), "4K 1spp", t(3840,2160,1), "4K 4spp", t(3840,2160,4))
